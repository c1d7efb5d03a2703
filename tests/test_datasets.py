"""CPU tests of the training-data formats / bucketed batching mirror (transformertts_b200/data/datasets.py; reference:
data/datasets.py, data/metadata_readers.py, extract_durations.py:108-115)."""
from random import Random

import numpy as np
import pytest
import torch

from transformertts_b200.data import datasets as ds


def _tok(text):
    return [1 + (ord(c) % 100) for c in text]


@pytest.fixture()
def corpus(tmp_path):
    rng = np.random.default_rng(0)
    names, lines = [], []
    for d in ('mels', 'durations', 'pitch_char'):
        (tmp_path / d).mkdir()
    for i in range(57):
        name = f'utt{i:03d}'
        n_tok = int(rng.integers(5, 40))
        text = ''.join(chr(97 + int(c)) for c in rng.integers(0, 26, n_tok))
        if i % 9 == 0:
            text += '?' if i % 2 else '!'
        dur = rng.integers(0, 8, len(text)).astype(np.int32)
        mel = rng.normal(-5, 2, (int(dur.sum()), 80)).astype(np.float32)
        np.save(tmp_path / 'mels' / f'{name}.npy', mel)
        np.save(tmp_path / 'durations' / f'{name}.npy', dur)
        np.save(tmp_path / 'pitch_char' / f'{name}.npy', rng.normal(0, 1, len(text)).astype(np.float32))
        names.append(name)
        lines.append(f'{name}|{text}\n')
    (tmp_path / 'train_metafile.txt').write_text(''.join(lines), encoding='utf-8')
    (tmp_path / 'metadata.csv').write_text(''.join(f'{n}.wav|raw text|{l.split("|")[1]}' for n, l in zip(names, lines)), encoding='utf-8')
    return tmp_path, names


def test_metadata_readers_and_upsampling(corpus):
    root, names = corpus
    td, up = ds.post_processed_reader(root / 'train_metafile.txt')
    assert list(td) == names
    marked = [n for n in names if any(c in td[n] for c in '?!')]
    assert up == [n for n in marked for _ in range(10)]          # x10, in file order (metadata_readers.py:46-47)
    lj = ds.ljspeech(root / 'metadata.csv')
    assert list(lj) == names and lj[names[3]] == td[names[3]]   # '.wav' dropped, LAST column is the text
    r_train = ds.DataReader(root / 'train_metafile.txt', training=True, is_processed=True)
    r_valid = ds.DataReader(root / 'train_metafile.txt', training=False, is_processed=True)
    assert len(r_train.filenames) == len(names) + 10 * len(marked) and r_valid.filenames == names


def _reference_batches(lengths, order, boundaries, sizes, drop_remainder):
    """brute-force restatement of bucket_by_sequence_length on a fixed sample order -> list of index lists"""
    buckets = [[] for _ in sizes]
    out = []
    for i in order:
        b = sum(1 for x in boundaries if lengths[i] >= x)
        buckets[b].append(i)
        if len(buckets[b]) == sizes[b]:
            out.append(buckets[b])
            buckets[b] = []
    if not drop_remainder:
        out += [b for b in buckets if b]
    return out


@pytest.mark.parametrize('drop_remainder', [False, True])
def test_bucketing_padding_and_order(corpus, drop_remainder):
    root, names = corpus
    reader = ds.DataReader(root / 'train_metafile.txt', training=False, is_processed=True)
    pre = ds.TTSPreprocessor(80, _tok)
    data = ds.TTSDataset(reader, pre, root / 'mels', root / 'durations', root / 'pitch_char')
    boundaries, sizes = [60, 100, 140], [6, 5, 4, 3]
    dset = data.get_dataset(bucket_batch_sizes=sizes, bucket_boundaries=boundaries, shuffle=True, drop_remainder=drop_remainder,
                            pin_memory=False)
    lengths = {n: np.load(root / 'mels' / f'{n}.npy').shape[0] for n in names}
    order = names[:]
    Random(42).shuffle(order)                                    # datasets.py:241,289: one shuffle per pass, seed 42
    want = _reference_batches(lengths, order, boundaries, sizes, drop_remainder)
    got = list(dset.all_batches())
    assert [b['name'] for b in got] == want
    for b in got:
        n = len(b['name'])
        T = max(lengths[x] for x in b['name'])
        assert b['mel'].shape == (n, T, 80) and b['mel'].dtype == torch.float32
        assert b['tokens'].dtype == torch.int32 and b['durations'].dtype == torch.int32 and b['pitch'].dtype == torch.float32
        for i, x in enumerate(b['name']):
            mel = np.load(root / 'mels' / f'{x}.npy')
            assert np.array_equal(b['mel'][i, :mel.shape[0]].numpy(), mel) and float(b['mel'][i, mel.shape[0]:].abs().sum()) == 0.0
            dur = np.load(root / 'durations' / f'{x}.npy')
            assert np.array_equal(b['durations'][i, :len(dur)].numpy(), dur) and int(b['durations'][i, len(dur):].sum()) == 0
            assert b['tokens'][i, :len(dur)].tolist() == _tok(reader.text_dict[x])
            assert int(b['durations'][i].sum()) == mel.shape[0]     # what the length regulator relies on


def test_endless_iteration_reshuffles_every_pass(corpus):
    root, names = corpus
    reader = ds.DataReader(root / 'train_metafile.txt', is_processed=True)
    data = ds.TTSDataset(reader, ds.TTSPreprocessor(80, _tok), root / 'mels', root / 'durations', root / 'pitch_char')
    dset = data.get_dataset(bucket_batch_sizes=[4, 4], bucket_boundaries=[100], shuffle=True, drop_remainder=False, pin_memory=False)
    per_pass = len(list(data.get_dataset(bucket_batch_sizes=[4, 4], bucket_boundaries=[100], drop_remainder=False, pin_memory=False).all_batches()))
    first = [dset.next_batch()['name'] for _ in range(per_pass)]
    second = [dset.next_batch()['name'] for _ in range(per_pass)]
    flat = lambda bs: sorted(x for b in bs for x in b)
    assert flat(first) == sorted(names) and flat(second) == sorted(names)
    assert first != second


def test_data_parallel_slices_are_disjoint_and_complete(corpus):
    root, names = corpus
    reader = ds.DataReader(root / 'train_metafile.txt', is_processed=True)
    data = ds.TTSDataset(reader, ds.TTSPreprocessor(80, _tok), root / 'mels', root / 'durations', root / 'pitch_char')
    kw = dict(bucket_batch_sizes=[6, 4], bucket_boundaries=[100], shuffle=True, drop_remainder=True, pin_memory=False)
    whole = list(data.get_dataset(**kw).all_batches())
    parts = [list(data.get_dataset(rank=r, world_size=2, **kw).all_batches()) for r in range(2)]
    assert len(parts[0]) == len(parts[1]) == len(whole)
    for w, a, b in zip(whole, parts[0], parts[1]):
        assert a['name'] == w['name'][0::2] and b['name'] == w['name'][1::2]


def test_aligner_samples_have_start_end_vectors_and_stop_targets(corpus):
    root, names = corpus
    reader = ds.DataReader(root / 'train_metafile.txt', is_processed=True)
    pre = ds.AlignerPreprocessor(80, 0.5, -0.5, _tok)
    data = ds.AlignerDataset(reader, pre, root / 'mels')
    b = next(iter(data.get_dataset(bucket_batch_sizes=[3, 3], bucket_boundaries=[100], shuffle=False, pin_memory=False).all_batches()))
    for i, x in enumerate(b['name']):
        mel = np.load(root / 'mels' / f'{x}.npy')
        T = mel.shape[0] + 2
        assert torch.all(b['mel'][i, 0] == 0.5) and torch.all(b['mel'][i, T - 1] == -0.5)
        assert np.array_equal(b['mel'][i, 1:T - 1].numpy(), mel)
        assert b['stop_prob'][i, :T].tolist() == [1] * (T - 1) + [2] and int(b['stop_prob'][i, T:].sum()) == 0


def test_pitch_per_char_matches_definition():
    pitch = np.array([0.0, 1.0, 3.0, 0.0, 0.0, 50.0, 2.0, 4.0], dtype=np.float32)
    dur = np.array([3, 2, 0, 3], dtype=np.int32)
    got = ds.pitch_per_char(pitch, dur, mel_len=8, pitch_mean=100.0, pitch_std=10.0)
    # char 0: frames 0-2 -> non-zero {1,3} -> 2.0 ; char 1: frames 3-4 all zero -> 0 ; char 2: no frames -> 0 ;
    # char 3: frames 5-7 -> 50 de-normalises to 600 Hz (>= 400, dropped) -> mean{2,4} = 3
    assert np.allclose(got, [2.0, 0.0, 0.0, 3.0])


def test_prefetch_loader_yields_the_same_stream(corpus):
    root, names = corpus
    reader = ds.DataReader(root / 'train_metafile.txt', is_processed=True)
    data = ds.TTSDataset(reader, ds.TTSPreprocessor(80, _tok), root / 'mels', root / 'durations', root / 'pitch_char')
    kw = dict(bucket_batch_sizes=[5, 5], bucket_boundaries=[100], shuffle=True, drop_remainder=True, pin_memory=False)
    direct = data.get_dataset(**kw)
    loader = ds.PrefetchLoader(data.get_dataset(**kw), prefetch=3)
    try:
        for _ in range(12):
            a, b = direct.next_batch(), loader.next()
            assert a['name'] == b['name'] and torch.equal(a['mel'], b['mel'])
    finally:
        loader.close()


def test_data_parallel_shards_have_equal_shapes_and_cover_each_global_batch(tmp_path):
    """Data-parallel batching: bucket sizes must be multiples of the world size, every rank walks the same stream
    of global batches, shards are padded to the GLOBAL lengths (equal shapes -> the mean of the shard losses is the global
    loss) and the remainder flush is cut identically on every rank."""
    rng = np.random.default_rng(3)
    samples = [f's{i}' for i in range(37)]
    lens = {s: int(rng.integers(5, 60)) for s in samples}

    def prep(name):
        n = lens[name]
        return np.full((n, 2), float(n), dtype=np.float32), list(range(1, 1 + n // 4 + 1)), name

    def make(rank, world, sizes=(4, 2)):
        return ds.Dataset(samples, prep, lambda m, *_: m.shape[0], ('mel', 'tokens', 'name'), (np.float32, np.int32, None), [30], list(sizes),
                          shuffle=True, drop_remainder=False, seed=1, pin_memory=False, rank=rank, world_size=world)

    with pytest.raises(ValueError):
        make(0, 2, sizes=(5, 2))
    assert ds.round_batch_sizes([64, 42, 25, 1], 8) == [64, 40, 24, 8]
    single = list(make(0, 1).all_batches())
    r0, r1 = list(make(0, 2).all_batches()), list(make(1, 2).all_batches())
    assert len(r0) == len(r1)
    seen = []
    for a, b in zip(r0, r1):
        assert a['mel'].shape == b['mel'].shape and a['tokens'].shape == b['tokens'].shape      # equal shapes on both ranks
        assert len(a['name']) == len(b['name'])
        seen += a['name'] + b['name']
    full = [n for bt in single for n in bt['name']]
    assert set(seen) <= set(full) and len(seen) >= len(full) - 2                                 # at most world_size - 1 rows cut per bucket tail


def test_config_manager_paths_and_latest_checkpoint(tmp_path):
    import yaml
    from pathlib import Path
    from transformertts_b200.utils.training_config_manager import TrainingConfigManager
    root = Path(__file__).resolve().parent.parent
    raw = yaml.safe_load((root / 'config' / 'training_config.yaml').read_text())
    raw['paths']['log_directory'] = str(tmp_path / 'logs')
    raw['paths']['train_data_directory'] = str(tmp_path / 'data')
    cfg = tmp_path / 'c.yaml'
    cfg.write_text(yaml.safe_dump(raw))
    cm = TrainingConfigManager(str(cfg))
    # the reference's naming scheme (utils/training_config_manager.py:23-44)
    assert cm.weights_dir == tmp_path / 'logs' / 'ljspeech' / 'tts_swap_conv_dims.alinger_extralayer_layernorm' / 'weights'
    assert cm.mel_dir.name == 'mels.MelGAN_default' and cm.data_dir.name == 'data.ljspeech'
    assert cm.duration_dir.name == 'durations.alinger_extralayer_layernorm.Stress_NoBreathing.MelGAN_default'
    assert cm.train_metadata_path.name == 'train_metadata.Stress_NoBreathing.txt'
    cm.create_remove_dirs()
    assert cm.latest_checkpoint() is None
    for name in ('step_5', 'step_20', 'step_100'):
        (cm.weights_dir / name).mkdir()
        (cm.weights_dir / name / 'optimizer.pt').write_bytes(b'x')
    (cm.weights_dir / 'step_300').mkdir()                        # weights only: cannot resume from it
    assert cm.latest_checkpoint().name == 'step_100'
    (cm.weights_dir / 'latest').mkdir()
    (cm.weights_dir / 'latest' / 'optimizer.pt').write_bytes(b'x')
    assert cm.latest_checkpoint().name == 'latest'
    cm.create_remove_dirs(clear_weights=True)
    assert cm.latest_checkpoint() is None
