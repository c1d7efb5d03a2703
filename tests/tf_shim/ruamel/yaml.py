"""Import stub for ruamel.yaml backed by PyYAML (only YAML().load / .dump are used by the reference)."""
import yaml as _yaml


class YAML:
    def load(self, stream):
        return _yaml.safe_load(stream)

    def dump(self, data, stream):
        _yaml.safe_dump(data, stream)
