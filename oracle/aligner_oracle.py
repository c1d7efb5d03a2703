"""CPU oracle for the Aligner teacher-forced step (BASELINE config C5).

TEST INFRASTRUCTURE ONLY -- same rules as oracle/forward_oracle.py: nothing under ``oracle/`` is imported by the
product package; only tests/, __graft_entry__.smoke() and bench.py's CPU legs use it, as the checker.

PARITY STATUS: pinned to the reference's own code like oracle/forward_oracle.py -- tests/test_reference_shim.py runs the
unmodified reference ``Aligner`` (model/models.py:15-341) on tests/tf_shim and compares its validation step (mel, stop logits,
all attention maps, the three losses; r = 1 and 2, diagonal losses on) with this file; tests/golden/aligner_small.npz is
written by that reference-code run.  Pinned against the reference's own known answers in addition:
  * the stop-token cross entropy -- tests/test_loss.py:12-24 of the reference (2.3705523014068604 with scaling 5,
    0.7679619193077087 with scaling 1 / masked_crossentropy) -- see tests/test_aligner_oracle.py
  * the look-ahead mask against ``torch.triu`` and the block structure against an independent
    ``torch.nn.functional.scaled_dot_product_attention`` implementation -- tests/test_aligner_oracle.py

Every function cites the reference lines it restates.  Keras layouts are kept (Dense kernel (in, out)).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import forward_oracle as fo

Tensor = torch.Tensor
ALIGNER_VOCAB = 129  # 126 symbols + pad + start + end ids (data/text/tokenizer.py:17-25 with add_start_end=True)


# ----------------------------------------------------------------------------------------
# model/transformer_utils.py
# ----------------------------------------------------------------------------------------
def create_look_ahead_mask(size: int) -> Tensor:
    """transformer_utils.py:35-37: 1 - band_part(ones, -1, 0) -> 1 strictly above the diagonal."""
    return 1.0 - torch.tril(torch.ones(size, size))


# ----------------------------------------------------------------------------------------
# model/layers.py
# ----------------------------------------------------------------------------------------
def decoder_prenet(p, x: Tensor, rate: float, training: bool, gen) -> Tensor:
    """DecoderPrenet layers.py:420-443: relu Dense -> dropout -> relu Dense -> dropout (dropout only when training)."""
    x = fo.dense(x, p['prenet.d1.w'], p['prenet.d1.b'], 'relu')
    x = fo.dropout(x, rate, training, gen)
    x = fo.dense(x, p['prenet.d2.w'], p['prenet.d2.b'], 'relu')
    x = fo.dropout(x, rate, training, gen)
    return x


def cross_attention_dense_block(p, pre: str, x: Tensor, enc_output: Tensor, look_ahead_mask: Tensor, padding_mask: Tensor,
                                num_heads: int, rate: float, training: bool, gen):
    """CrossAttentionDenseBlock layers.py:330-349 = SelfAttentionResNorm (:198-211) -> CrossAttentionResnorm (:315-327)
    -> FFNResNorm (:82-102).  No row masking anywhere in this block."""
    a1, w1 = fo.multi_head_attention(p, pre + 'sa.', x, x, x, look_ahead_mask, num_heads, rate, training, gen)
    attn1 = fo.layer_norm(a1 + x, p[pre + 'sa.ln.gamma'], p[pre + 'sa.ln.beta'])
    a2, w2 = fo.multi_head_attention(p, pre + 'ca.', enc_output, enc_output, attn1, padding_mask, num_heads, rate, training, gen)
    attn2 = fo.layer_norm(a2 + attn1, p[pre + 'ca.ln.gamma'], p[pre + 'ca.ln.beta'])
    h = fo.dense(attn2, p[pre + 'ffn1.w'], p[pre + 'ffn1.b'], 'relu')
    h = fo.dense(h, p[pre + 'ffn2.w'], p[pre + 'ffn2.b'])
    h = fo.dropout(h, rate, training, gen)
    out = fo.layer_norm(h + attn2, p[pre + 'ln2.gamma'], p[pre + 'ln2.beta'])
    return out, w1, w2


def cross_attention_blocks(p, cfg: dict, inputs: Tensor, enc_output: Tensor, decoder_padding_mask: Tensor,
                           encoder_padding_mask: Tensor, reduction_factor: int, training: bool, gen):
    """CrossAttentionBlocks layers.py:381-417: LN(inputs) + scalar * PE[:, :T*r:r] -> dropout -> blocks."""
    T = inputs.shape[1]
    r = int(reduction_factor)
    x = fo.layer_norm(inputs, p['decoder.ln.gamma'], p['decoder.ln.beta'])
    pe = fo.positional_encoding(int(cfg['decoder_max_position_encoding']), int(cfg['decoder_model_dimension']))
    x = x + p['decoder.pos_scalar'] * pe[:, :T * r:r, :].to(x.dtype)
    x = fo.dropout(x, float(cfg['dropout_rate']), training, gen)
    weights = {}
    heads = list(cfg['decoder_num_heads'])
    for i, nh in enumerate(heads):
        x, _, w = cross_attention_dense_block(p, f'decoder.b{i}.', x, enc_output, decoder_padding_mask, encoder_padding_mask,
                                              nh, float(cfg['dropout_rate']), training, gen)
        key = 'Decoder_LastBlock_CrossAttention' if i == len(heads) - 1 else f'Decoder_DenseBlock{i + 1}_CrossAttention'
        weights[key] = w
    return x, weights


def aligner_call(p: Dict[str, Tensor], cfg: dict, inputs: Tensor, targets: Tensor, r: int = 1, training: bool = False,
                 gen: Optional[torch.Generator] = None) -> dict:
    """Aligner.call models.py:294-298 = _call_encoder (:127-133) + _call_decoder (:135-154)."""
    rate = float(cfg['dropout_rate'])
    padding_mask = fo.create_encoder_padding_mask(inputs)
    enc_in = p['embedding'][inputs.long()]
    enc_stack = {'num_heads': list(cfg['encoder_num_heads']), 'dense_blocks': len(cfg['encoder_num_heads']), 'dropout': rate,
                 'pe': fo.positional_encoding(int(cfg['encoder_max_position_encoding']), int(cfg['encoder_model_dimension']))}
    enc_output, enc_attn = fo.self_attention_blocks(p, 'encoder', enc_stack, enc_in, padding_mask, training, gen)
    # ---- decoder
    dec_target_padding_mask = fo.create_mel_padding_mask(targets)                       # (B,1,1,T)
    look_ahead = create_look_ahead_mask(targets.shape[1]).to(targets.dtype)             # (T,T)
    combined = torch.maximum(dec_target_padding_mask, look_ahead)                       # (B,1,T,T)
    dec_input = decoder_prenet(p, targets, float(cfg['decoder_prenet_dropout']), training, gen)
    dec_output, dec_attn = cross_attention_blocks(p, cfg, dec_input, enc_output, combined, padding_mask, r, training, gen)
    mel_ch = int(cfg['mel_channels'])
    out_proj = fo.dense(dec_output, p['final_proj.w'], p['final_proj.b'])[:, :, :r * mel_ch]
    B, T = out_proj.shape[:2]
    linear = out_proj.reshape(B, T * r, mel_ch)
    stop = fo.dense(linear, p['postnet.stop.w'], p['postnet.stop.b'])   # Postnet layers.py:446-460
    mel = fo.dense(linear, p['postnet.mel.w'], p['postnet.mel.b'])
    return {'mel': mel, 'stop_prob': stop, 'decoder_attention': dec_attn, 'decoder_output': dec_output, 'linear': linear,
            'mel_mask': dec_target_padding_mask, 'encoder_attention': enc_attn, 'text_mask': padding_mask}


def aligner_predict(p: Dict[str, Tensor], cfg: dict, inp: Tensor, start_value: float, max_length: int = 1000, r: int = 1,
                    stop_prob_index: int = 2) -> dict:
    """Aligner.predict models.py:271-292 (encode=False): autoregressive decoding of ONE token row.  The decoder is re-run on
    the whole prefix every iteration (as the reference does); the prefix grows by the LAST predicted frame (one per iteration,
    also for r > 1, :280), the returned mel by the last r frames (:281-282); decoding stops when the arg-max of the last stop
    distribution is `stop_prob_index` (:287) or after max_length // r + 1 iterations."""
    mel_ch = int(cfg['mel_channels'])
    inp = inp.reshape(1, -1)
    output = torch.full((1, 1, mel_ch), float(start_value), dtype=p['embedding'].dtype)
    output_concat = output.clone()
    out = {}
    for _ in range(int(max_length // r) + 1):
        mo = aligner_call(p, cfg, inp, output, r=r, training=False)
        output = torch.cat([output, mo['mel'][:1, -1:, :]], dim=-2)
        output_concat = torch.cat([output_concat, mo['mel'][:1, -r:, :]], dim=-2)
        out = {'mel': output_concat[0, 1:, :], 'decoder_attention': mo['decoder_attention'],
               'encoder_attention': mo['encoder_attention'], 'stop_prob': mo['stop_prob']}
        if int(torch.argmax(mo['stop_prob'][:, -1], dim=-1)) == stop_prob_index:
            break
    return out


# ----------------------------------------------------------------------------------------
# utils/losses.py, utils/metrics.py
# ----------------------------------------------------------------------------------------
def new_scaled_crossentropy(targets: Tensor, logits: Tensor, index: int = 2, scaling: float = 1.0) -> Tensor:
    """utils/losses.py:4-21.  Keras SparseCategoricalCrossentropy(from_logits) with sample_weight and the default
    SUM_OVER_BATCH_SIZE reduction: sum(ce * weight) / number of elements (NOT / sum of weights)."""
    ce = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(), targets.reshape(-1).long(), reduction='none')
    t = targets.reshape(-1)
    w = (t != 0).float() + (t == index).float() * (scaling - 1.0)
    return (ce * w).sum() / ce.numel()


def diagonal_mask(mel_len: int, phon_len: int, padded_shape) -> Tensor:
    """utils/metrics.py:59-70: |i/max_n - j/max_m| on the valid (mel_len, phon_len) corner, zero elsewhere."""
    max_m = min(int(mel_len), int(padded_shape[0]))
    max_n = int(phon_len)
    i = torch.arange(max_n, dtype=torch.float64)[None, :].expand(max_m, max_n) / max_n
    j = torch.arange(max_m, dtype=torch.float64)[:, None].expand(max_m, max_n) / max_m
    d = torch.sqrt((i - j) ** 2)
    out = torch.zeros(tuple(padded_shape), dtype=torch.float64)
    out[:max_m, :max_n] = d
    return out.float()


def batch_diagonal_mask(att: Tensor, mel_len: Tensor, phon_len: Tensor) -> Tensor:
    """utils/metrics.py:47-57 -> (B,1,Tq,Tk)."""
    B, _, M, N = att.shape
    return torch.stack([diagonal_mask(int(mel_len[b]), int(phon_len[b]), (M, N)) for b in range(B)])[:, None]


def gta_forward(p, cfg, inp: Tensor, tar: Tensor, stop_prob: Tensor, r: int = 1, stop_scaling: float = 8.0,
                force_encoder_diagonal: bool = False, force_decoder_diagonal: bool = False, training: bool = False, gen=None) -> dict:
    """Aligner._gta_forward models.py:168-210 (forward + losses; the tape/optimizer part is out of this row's scope)."""
    tar_inp = tar[:, :-1]
    tar_real = tar[:, 1:]
    tar_stop = stop_prob[:, 1:]
    mel_len = tar_inp.shape[1]
    tar_mel = tar_inp[:, 0::r, :]
    out = aligner_call(p, cfg, inp, tar_mel, r=r, training=training, gen=gen)
    mel_loss = fo.masked_mean_absolute_error(tar_real, out['mel'][:, :mel_len, :])
    stop_loss = new_scaled_crossentropy(tar_stop, out['stop_prob'][:, :mel_len, :], index=2, scaling=stop_scaling)
    loss = mel_loss + stop_loss  # loss_weights [1., 1.] (models.py:223)
    phon_len = (1.0 - out['text_mask'][:, 0, 0, :]).sum(dim=1)
    d_loss = torch.tensor(0.0)
    norm = 1.0
    if force_decoder_diagonal:
        m_len = (1.0 - out['mel_mask'][:, 0, 0, :]).sum(dim=1)
        keys = list(out['decoder_attention'].keys())
        dmask = batch_diagonal_mask(out['decoder_attention'][keys[0]], m_len, phon_len)
        for k in keys:
            d_loss = d_loss + (out['decoder_attention'][k] * dmask).sum(dim=(-2, -1)).mean() / 10.0
        norm += len(keys)
    if force_encoder_diagonal:
        keys = list(out['encoder_attention'].keys())
        dmask = batch_diagonal_mask(out['encoder_attention'][keys[0]], phon_len, phon_len)
        for k in keys:
            d_loss = d_loss + (out['encoder_attention'][k] * dmask).sum(dim=(-2, -1)).mean() / 10.0
        norm += len(keys)
    d_loss = d_loss / norm
    loss = loss + d_loss
    out['loss'] = loss
    out['losses'] = {'mel': mel_loss, 'stop_prob': stop_loss, 'diag_loss': d_loss}
    return out


# ----------------------------------------------------------------------------------------
# configs, parameters, inputs
# ----------------------------------------------------------------------------------------
ALIGNER_CONFIGS = {
    # config/training_config.yaml:58-69 (aligner_settings as shipped) + :17-19 (mel start/end values)
    'A5': dict(encoder_model_dimension=256, decoder_model_dimension=256, encoder_num_heads=[4, 4, 4, 4],
               decoder_num_heads=[4, 4, 4, 4, 1], encoder_feed_forward_dimension=512, decoder_feed_forward_dimension=512,
               encoder_prenet_dimension=256, decoder_prenet_dimension=256, encoder_max_position_encoding=10000,
               decoder_max_position_encoding=10000, dropout_rate=0.1, decoder_prenet_dropout=0.1, mel_channels=80,
               mel_start_value=0.5, mel_end_value=-0.5, max_r=10, stop_loss_scaling=8, vocab_size=ALIGNER_VOCAB,
               phoneme_language='en-us', with_stress=True, model_breathing=False),
    # plumbing-size variant: 2 encoder blocks, 2 decoder blocks (last with one head), d=128
    'A-small': dict(encoder_model_dimension=128, decoder_model_dimension=128, encoder_num_heads=[2, 2],
                    decoder_num_heads=[2, 1], encoder_feed_forward_dimension=256, decoder_feed_forward_dimension=256,
                    encoder_prenet_dimension=128, decoder_prenet_dimension=192, encoder_max_position_encoding=2000,
                    decoder_max_position_encoding=4000, dropout_rate=0.1, decoder_prenet_dropout=0.1, mel_channels=80,
                    mel_start_value=0.5, mel_end_value=-0.5, max_r=4, stop_loss_scaling=8, vocab_size=ALIGNER_VOCAB,
                    phoneme_language='en-us', with_stress=True, model_breathing=False),
}


def init_aligner_params(cfg: dict, seed: int = 7, dtype=torch.float32) -> Dict[str, Tensor]:
    """Same recipe as forward_oracle.init_params (SURVEY 8(c) weights row): Glorot kernels, biases 0.05*N, LN 1+0.1*N / 0.1*N,
    embedding U(-.05,.05), pos_scalar 0.8 (encoder) / 1.1 (decoder)."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    p: Dict[str, Tensor] = {}
    d_enc, d_dec = int(cfg['encoder_model_dimension']), int(cfg['decoder_model_dimension'])
    mel = int(cfg['mel_channels'])

    def ln(prefix, n):
        p[prefix + '.gamma'] = 1.0 + 0.1 * torch.randn(n, generator=g)
        p[prefix + '.beta'] = 0.1 * torch.randn(n, generator=g)

    def lin(prefix, fin, fout):
        p[prefix + '.w'] = fo._glorot(g, (fin, fout), fin, fout)
        p[prefix + '.b'] = 0.05 * torch.randn(fout, generator=g)

    def mha(prefix, d_q, d_kv, d):
        lin(prefix + 'wq', d_q, d)
        lin(prefix + 'wk', d_kv, d)
        lin(prefix + 'wv', d_kv, d)
        lin(prefix + 'wo', d_q + d, d)

    assert int(cfg['encoder_prenet_dimension']) == d_enc, 'the embedding feeds the encoder blocks directly (models.py:53-65)'
    p['embedding'] = (torch.rand((int(cfg['vocab_size']), d_enc), generator=g) * 2 - 1) * 0.05
    ln('encoder.ln', d_enc)
    p['encoder.pos_scalar'] = torch.tensor(0.8)
    for i, _ in enumerate(cfg['encoder_num_heads']):
        pre = f'encoder.b{i}.'
        mha(pre, d_enc, d_enc, d_enc)
        ln(pre + 'ln1', d_enc)
        lin(pre + 'ffn1', d_enc, int(cfg['encoder_feed_forward_dimension']))
        lin(pre + 'ffn2', int(cfg['encoder_feed_forward_dimension']), d_enc)
        ln(pre + 'ln2', d_enc)
    lin('prenet.d1', mel, int(cfg['decoder_prenet_dimension']))
    lin('prenet.d2', int(cfg['decoder_prenet_dimension']), d_dec)
    ln('decoder.ln', d_dec)
    p['decoder.pos_scalar'] = torch.tensor(1.1)
    for i, _ in enumerate(cfg['decoder_num_heads']):
        pre = f'decoder.b{i}.'
        mha(pre + 'sa.', d_dec, d_dec, d_dec)
        ln(pre + 'sa.ln', d_dec)
        mha(pre + 'ca.', d_dec, d_enc, d_dec)
        ln(pre + 'ca.ln', d_dec)
        lin(pre + 'ffn1', d_dec, int(cfg['decoder_feed_forward_dimension']))
        lin(pre + 'ffn2', int(cfg['decoder_feed_forward_dimension']), d_dec)
        ln(pre + 'ln2', d_dec)
    lin('final_proj', d_dec, mel * int(cfg['max_r']))
    lin('postnet.stop', mel, 3)
    lin('postnet.mel', mel, mel)
    return {k: v.to(dtype) for k, v in p.items()}


def make_aligner_inputs(cfg: dict, B: int, Tp: int, Tm: int, seed: int = 500, ragged: bool = True):
    """tokens with start/end ids, mel with start/end vectors, stop targets 1..1,2 then 0 padding
    (data/datasets.py:85-95 of the reference: start vector + mel + end vector; stop = ones, last = 2)."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    vocab = int(cfg['vocab_size'])
    mel_ch = int(cfg['mel_channels'])
    start_id, end_id = vocab - 2, vocab - 1
    tokens = torch.zeros((B, Tp), dtype=torch.int32)
    mel = torch.zeros((B, Tm, mel_ch))
    stop = torch.zeros((B, Tm), dtype=torch.int32)
    for b in range(B):
        tp = Tp if (b == 0 or not ragged) else int(torch.randint(max(3, Tp // 2), Tp + 1, (1,), generator=g))
        tm = Tm if (b == 0 or not ragged) else int(torch.randint(max(4, Tm // 2), Tm + 1, (1,), generator=g))
        tokens[b, 0] = start_id
        tokens[b, 1:tp - 1] = torch.randint(1, vocab - 2, (tp - 2,), generator=g, dtype=torch.int32)
        tokens[b, tp - 1] = end_id
        body = torch.clamp(-5.0 + 2.0 * torch.randn((tm - 2, mel_ch), generator=g), -11.5, 2.0)
        mel[b, 0] = float(cfg['mel_start_value'])
        mel[b, 1:tm - 1] = body
        mel[b, tm - 1] = float(cfg['mel_end_value'])
        stop[b, :tm - 1] = 1
        stop[b, tm - 1] = 2
    return tokens, mel, stop


def loss_and_grads(p: Dict[str, Tensor], cfg: dict, inp: Tensor, tar: Tensor, stop_prob: Tensor, r: int = 1,
                   stop_scaling: float = 8.0, force_encoder_diagonal: bool = False, force_decoder_diagonal: bool = False):
    """Forward (dropout off) + backward with torch autograd on the restated graph: the oracle of the hand-written
    Aligner backward pass (models.py:212-216).  Returns (outputs, {name: grad})."""
    q = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    out = gta_forward(q, cfg, inp, tar, stop_prob, r=r, stop_scaling=stop_scaling, force_encoder_diagonal=force_encoder_diagonal,
                      force_decoder_diagonal=force_decoder_diagonal, training=False)
    out['loss'].backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in q.items()}
    return out, grads
