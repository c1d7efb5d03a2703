"""Keras ``model_weights.hdf5`` interop without h5py -- the pure-python HDF5 subset reader / writer
(transformertts_b200/utils/hdf5_lite.py) and the Keras layer/weight ORDER map (transformertts_b200/model/hdf5_weights.py)."""
import struct

import numpy as np
import pytest
import torch

from oracle import forward_oracle as fo
from transformertts_b200.model import hdf5_weights as hw
from transformertts_b200.utils import hdf5_lite as h5


def test_hdf5_lite_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    root = h5.Group()
    root.attrs['layer_names'] = np.array([b'Embedding', b'Encoder', b'dense_17'])
    root.attrs['backend'] = np.array(b'tensorflow')
    root.attrs['answer'] = np.array([1, 2, 3], dtype=np.int32)
    root.set_dataset('Encoder/Encoder/Variable:0', np.float32(0.8))                       # scalar dataset
    root.set_dataset('Encoder/Encoder/blk/kernel:0', rng.normal(size=(3, 5, 7)).astype(np.float32))
    root.set_dataset('Encoder/Encoder/blk/bias:0', rng.normal(size=(7,)).astype(np.float32))
    root.set_dataset('ints', np.arange(12, dtype=np.int64).reshape(3, 4))
    root.set_dataset('halves', rng.normal(size=(4,)).astype(np.float16))
    root.set_dataset('doubles', rng.normal(size=(2, 2)))
    root.require_group('expand').attrs['weight_names'] = np.zeros((0,), dtype='S1')       # a layer without weights
    many = root.require_group('many')
    for i in range(100):                                                                   # > 8 links in one group
        many.set_dataset(f'w{i:03d}', np.full((2,), i, dtype=np.float32))
    path = tmp_path / 'x.hdf5'
    h5.write_hdf5(path, root)
    raw = path.read_bytes()
    assert raw[:8] == b'\x89HDF\r\n\x1a\n' and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    assert struct.unpack_from('<Q', raw, 40)[0] == len(raw)                               # end-of-file address
    back = h5.read_hdf5(path)
    assert [n.decode() for n in back.attrs['layer_names']] == ['Embedding', 'Encoder', 'dense_17']
    assert back.attrs['backend'].item() == b'tensorflow' and back.attrs['answer'].tolist() == [1, 2, 3]
    assert back['Encoder/Encoder/Variable:0'].shape == () and back['Encoder/Encoder/Variable:0'] == np.float32(0.8)
    for k in ('Encoder/Encoder/blk/kernel:0', 'Encoder/Encoder/blk/bias:0', 'ints', 'halves', 'doubles'):
        assert back[k].dtype == root[k].dtype and np.array_equal(back[k], root[k]), k
    assert back['expand'].attrs['weight_names'].shape == (0,)
    assert sorted(back['many'].children) == [f'w{i:03d}' for i in range(100)]
    assert all(back['many'].children[f'w{i:03d}'][0] == i for i in range(100))


def test_reader_rejects_what_it_does_not_implement(tmp_path):
    p = tmp_path / 'bad.hdf5'
    p.write_bytes(b'not an hdf5 file at all')
    with pytest.raises(h5.Hdf5Error):
        h5.read_hdf5(p)
    root = h5.Group()
    root.set_dataset('x', np.zeros(3, dtype=np.float32))
    h5.write_hdf5(p, root)
    raw = bytearray(p.read_bytes())
    raw[8] = 2                                                                             # superblock version 2
    p.write_bytes(bytes(raw))
    with pytest.raises(h5.Hdf5Error):
        h5.read_hdf5(p)


@pytest.mark.parametrize('cfg_name', ['C1', 'LJ256', 'LJ256-dense'])
def test_keras_weight_file_round_trip(tmp_path, cfg_name):
    from transformertts_b200.model.models import ForwardTransformer
    cfg = fo.CONFIGS[cfg_name]
    m = ForwardTransformer(**cfg, device='cpu')
    p = fo.init_params(cfg, seed=7)
    m.set_weights(p)
    hw.save_keras_hdf5(m, tmp_path / 'model_weights.hdf5')
    tree = h5.read_hdf5(tmp_path / 'model_weights.hdf5')
    names = [n.decode() for n in tree.attrs['layer_names']]
    assert names == ['Embedding', 'Encoder', 'dur_pred', 'expand', 'pitch_pred', 'dense', 'Decoder', 'dense_1']
    assert all(n.decode().startswith('Encoder/') for n in tree['Encoder'].attrs['weight_names'])
    m2 = ForwardTransformer(**cfg, device='cpu', seed=1)
    got = hw.load_keras_hdf5(m2, tmp_path / 'model_weights.hdf5')
    assert set(got) == set(p)
    for k in p:
        assert torch.equal(got[k], p[k]), k
    # a file of another architecture is refused by the shape checks (the loader matches by order, like Keras)
    other = ForwardTransformer(**fo.CONFIGS['REF384'], device='cpu')
    with pytest.raises(h5.Hdf5Error):
        hw.load_keras_hdf5(other, tmp_path / 'model_weights.hdf5')
