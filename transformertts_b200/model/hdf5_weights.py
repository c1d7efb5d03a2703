"""Keras ``model_weights.hdf5`` <-> flat parameter dictionary (reference: ForwardTransformer.save_model / load_model,
model/models.py:600-638: ``self.save_weights(path / 'model_weights.hdf5')`` / ``model.load_weights(...)``; published weights
``bdf06b9_ljspeech_step_*.zip``, model/factory.py:10-19).

Keras' HDF5 weight format (``save_weights_to_hdf5_group``): root attribute ``layer_names`` lists ``model.layers`` in
attribute-assignment order; each layer is a group whose ``weight_names`` attribute lists ``layer.weights`` in order (a
layer's own variables first, then its tracked sub-layers in the order they were assigned in ``__init__``), each weight a
contiguous float32 dataset at ``<group>/<weight name>``.  ``load_weights`` (by_name=False) assigns BY ORDER, checking only
counts and shapes -- the auto-generated names (``dense_17/kernel:0``) depend on TensorFlow's global name counters, so this
module maps by order too and verifies every shape.  The order below is read off the reference's constructors:

  ForwardTransformer.__init__ (models.py:381-422): encoder_prenet, encoder, dur_pred, expand, pitch_pred, pitch_embed, decoder, out
  SelfAttentionBlocks (layers.py:267-297): pos_encoding_scalar | dropout, encoder_SADB[...], encoder_SACB[...], layernorm
  SelfAttentionDenseBlock / ConvBlock (:214-264): sarn(mha(wq, wk, wv, dense), last_ln), then ffn(d1, d2, last_ln) or
      conv(convolutions[...], last_conv, normalization)
  StatPredictor (:463-485): conv_blocks(convolutions[...], last_conv, normalization[...]), linear

Tensor layouts need no conversion: the flat dictionary already uses Keras layouts (Dense (in,out), Conv1D (k,in,out)).
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, List, Tuple

import numpy as np
import torch

from ..utils.hdf5_lite import Group, Hdf5Error, read_hdf5, write_hdf5


def _stack_order(model, name: str) -> List[Tuple[str, str]]:
    st = model._stacks[name]
    lname = name.capitalize()
    out = [('Variable:0', f'{name}.pos_scalar')]
    for i, _ in enumerate(st['heads']):
        pre = f'{name}.b{i}.'
        dense = i < st['n_dense']
        blk = f'{lname}_SADB_{i}' if dense else f'{lname}_SACB_{i - st["n_dense"]}'
        mha = f'{blk}/self_attention_res_norm/multi_head_attention'
        for j, w in enumerate(('wq', 'wk', 'wv', 'wo')):
            out += [(f'{mha}/dense_{j}/kernel:0', pre + w + '.w'), (f'{mha}/dense_{j}/bias:0', pre + w + '.b')]
        out += [(f'{blk}/self_attention_res_norm/layer_normalization/gamma:0', pre + 'ln1.gamma'),
                (f'{blk}/self_attention_res_norm/layer_normalization/beta:0', pre + 'ln1.beta')]
        if dense:
            for j, w in enumerate(('ffn1', 'ffn2')):
                out += [(f'{blk}/ffn_res_norm/dense_{j}/kernel:0', pre + w + '.w'), (f'{blk}/ffn_res_norm/dense_{j}/bias:0', pre + w + '.b')]
            sub = 'ffn_res_norm'
        else:
            for j in range(len(st['filters'])):
                out += [(f'{blk}/cnn_res_norm/conv1d_{j}/kernel:0', pre + f'conv{j}.w'), (f'{blk}/cnn_res_norm/conv1d_{j}/bias:0', pre + f'conv{j}.b')]
            sub = 'cnn_res_norm'
        out += [(f'{blk}/{sub}/layer_normalization/gamma:0', pre + 'ln2.gamma'), (f'{blk}/{sub}/layer_normalization/beta:0', pre + 'ln2.beta')]
    out += [('layer_normalization/gamma:0', f'{name}.ln.gamma'), ('layer_normalization/beta:0', f'{name}.ln.beta')]
    return out


def _predictor_order(model, name: str) -> List[Tuple[str, str]]:
    n = len(model.config['duration_conv_filters' if name == 'dur_pred' else 'pitch_conv_filters'])
    out = []
    for j in range(n):
        out += [(f'cnn_dropout/conv1d_{j}/kernel:0', f'{name}.conv{j}.w'), (f'cnn_dropout/conv1d_{j}/bias:0', f'{name}.conv{j}.b')]
    for j in range(n):
        out += [(f'cnn_dropout/layer_normalization_{j}/gamma:0', f'{name}.ln{j}.gamma'), (f'cnn_dropout/layer_normalization_{j}/beta:0', f'{name}.ln{j}.beta')]
    out += [('dense/kernel:0', f'{name}.out.w'), ('dense/bias:0', f'{name}.out.b')]
    return out


def _mha_order(blk: str, sub: str, pre: str) -> List[Tuple[str, str]]:
    out = []
    for j, w in enumerate(('wq', 'wk', 'wv', 'wo')):
        out += [(f'{blk}/{sub}/multi_head_attention/dense_{j}/kernel:0', pre + w + '.w'),
                (f'{blk}/{sub}/multi_head_attention/dense_{j}/bias:0', pre + w + '.b')]
    return out


def _aligner_order(model) -> List[Tuple[str, List[Tuple[str, str]]]]:
    """Aligner.__init__ (models.py:53-63): encoder_prenet, encoder, decoder_prenet, decoder, final_proj_mel, decoder_postnet;
    CrossAttentionBlocks (layers.py:381-400): pos_encoding_scalar | dropout, CADB[...], last_CADB, layernorm;
    CrossAttentionDenseBlock (:330-341): sarn(mha, last_ln), carn(mha, layernorm), ffn(d1, d2, last_ln).
    DecoderPrenet also owns the NON-trainable dropout-rate variable (layers.py:432), which Keras lists after the trainable
    weights: it has no parameter here (flat name None; written from the config, ignored on load)."""
    dec = [('Variable:0', 'decoder.pos_scalar')]
    n = len(model._stacks['decoder']['heads'])
    for i in range(n):
        pre = f'decoder.b{i}.'
        blk = f'Decoder_CADB_{i}' if i < n - 1 else 'Decoder_CADB_last'
        dec += _mha_order(blk, 'self_attention_res_norm', pre + 'sa.')
        dec += [(f'{blk}/self_attention_res_norm/layer_normalization/gamma:0', pre + 'sa.ln.gamma'),
                (f'{blk}/self_attention_res_norm/layer_normalization/beta:0', pre + 'sa.ln.beta')]
        dec += _mha_order(blk, 'cross_attention_resnorm', pre + 'ca.')
        dec += [(f'{blk}/cross_attention_resnorm/layer_normalization/gamma:0', pre + 'ca.ln.gamma'),
                (f'{blk}/cross_attention_resnorm/layer_normalization/beta:0', pre + 'ca.ln.beta')]
        for j, w in enumerate(('ffn1', 'ffn2')):
            dec += [(f'{blk}/ffn_res_norm/dense_{j}/kernel:0', pre + w + '.w'), (f'{blk}/ffn_res_norm/dense_{j}/bias:0', pre + w + '.b')]
        dec += [(f'{blk}/ffn_res_norm/layer_normalization/gamma:0', pre + 'ln2.gamma'), (f'{blk}/ffn_res_norm/layer_normalization/beta:0', pre + 'ln2.beta')]
    dec += [('layer_normalization/gamma:0', 'decoder.ln.gamma'), ('layer_normalization/beta:0', 'decoder.ln.beta')]
    return [
        ('Embedding', [('embeddings:0', 'embedding')]),
        ('Encoder', _stack_order(model, 'encoder')),
        ('DecoderPrenet', [('dense/kernel:0', 'prenet.d1.w'), ('dense/bias:0', 'prenet.d1.b'), ('dense_1/kernel:0', 'prenet.d2.w'),
                           ('dense_1/bias:0', 'prenet.d2.b'), ('Variable:0', None)]),
        ('Decoder', dec),
        ('FinalProj', [('kernel:0', 'final_proj.w'), ('bias:0', 'final_proj.b')]),
        ('Postnet', [('dense/kernel:0', 'postnet.stop.w'), ('dense/bias:0', 'postnet.stop.b'), ('dense_1/kernel:0', 'postnet.mel.w'),
                     ('dense_1/bias:0', 'postnet.mel.b')]),
    ]


def keras_layer_order(model) -> List[Tuple[str, List[Tuple[str, str]]]]:
    """[(Keras layer name, [(weight name inside the layer group, flat parameter name), ...]), ...] in ``model.layers`` order."""
    if hasattr(model, 'max_r'):
        return _aligner_order(model)
    return [
        ('Embedding', [('embeddings:0', 'embedding')]),
        ('Encoder', _stack_order(model, 'encoder')),
        ('dur_pred', _predictor_order(model, 'dur_pred')),
        ('expand', []),
        ('pitch_pred', _predictor_order(model, 'pitch_pred')),
        ('dense', [('kernel:0', 'pitch_embed.w'), ('bias:0', 'pitch_embed.b')]),
        ('Decoder', _stack_order(model, 'decoder')),
        ('dense_1', [('kernel:0', 'out.w'), ('bias:0', 'out.b')]),
    ]


def _names_attr(names: List[str]) -> np.ndarray:
    return np.array([n.encode('utf-8') for n in names], dtype='S') if names else np.zeros((0,), dtype='S1')


def save_keras_hdf5(model, path) -> None:
    """Write ``model``'s weights in Keras' HDF5 weight format (what the reference's ``save_weights`` produces)."""
    root = Group()
    order = keras_layer_order(model)
    root.attrs['layer_names'] = _names_attr([lname for lname, _ in order])
    root.attrs['backend'] = np.array(b'tensorflow')
    root.attrs['keras_version'] = np.array(b'2.4.0')
    for lname, weights in order:
        g = root.require_group(lname)
        full = [f'{lname}/{w}' for w, _ in weights]
        g.attrs['weight_names'] = _names_attr(full)
        for fname, (_, flat) in zip(full, weights):
            if flat is None:   # DecoderPrenet's non-trainable dropout rate
                g.set_dataset(fname, np.float32(model.config.get('decoder_prenet_dropout', 0.0)))
            else:
                g.set_dataset(fname, model.weights[flat].detach().float().cpu().numpy())
    write_hdf5(path, root)


def load_keras_hdf5(model, path) -> Dict[str, torch.Tensor]:
    """Read a Keras ``model_weights.hdf5`` of the reference's ForwardTransformer into the flat parameter dictionary.
    Layers and weights are matched BY ORDER (as Keras' own loader does); every shape is checked."""
    root = read_hdf5(Path(path))
    if 'layer_names' not in root.attrs and 'model_weights' in root.children:
        root = root.children['model_weights']      # a full-model .h5 keeps the same structure one level down
    if 'layer_names' not in root.attrs:
        raise Hdf5Error('no layer_names attribute: not a Keras weight file')
    layer_names = [n.decode('utf-8') for n in np.atleast_1d(root.attrs['layer_names'])]
    stored = []
    for lname in layer_names:
        g = root[lname]
        wn = g.attrs.get('weight_names')
        if wn is None:   # Keras splits very long lists into weight_names0, weight_names1, ...
            parts, k = [], 0
            while f'weight_names{k}' in g.attrs:
                parts.append(np.atleast_1d(g.attrs[f'weight_names{k}']))
                k += 1
            wn = np.concatenate(parts) if parts else np.zeros((0,), dtype='S1')
        names = [n.decode('utf-8') for n in np.atleast_1d(wn)]
        if names:
            stored.append((lname, [(n, g[n]) for n in names]))
    expected = [(lname, w) for lname, w in keras_layer_order(model) if w]
    if len(stored) != len(expected):
        raise Hdf5Error(f'file holds {len(stored)} layers with weights ({[s[0] for s in stored]}), the model has {len(expected)}')
    shapes = model._param_shapes()
    out = {}
    for (file_layer, file_w), (_, want) in zip(stored, expected):
        if len(file_w) != len(want):
            raise Hdf5Error(f'layer {file_layer}: {len(file_w)} weights in the file, {len(want)} expected')
        for (wname, arr), (_, flat) in zip(file_w, want):
            if flat is None:
                continue
            if tuple(arr.shape) != tuple(shapes[flat]):
                raise Hdf5Error(f'{file_layer}/{wname}: shape {tuple(arr.shape)} does not match {flat} {tuple(shapes[flat])}')
            out[flat] = torch.from_numpy(np.array(arr, dtype=np.float32, order='C'))
    return out
