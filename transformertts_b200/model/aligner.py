"""Aligner: host-side mirror of the reference's teacher-forced encoder-decoder (model/models.py:15-341) -- the model
that produces the attention maps durations are extracted from.

Built here: the teacher-forced forward (``call`` / ``_forward`` / ``_forward_encoder`` / ``_forward_decoder``) and the
validation step with its losses (``_val_step`` = ``_gta_forward(training=False)``, models.py:168-220), every layer
through libttsb.so; ``_train_step`` (teacher-forced forward with dropout in single-pass bf16, hand-written backward,
Keras-form Adam) lives in aligner_training.py; ``predict`` is the reference's autoregressive loop over the same decoder call.

Parameter names (flat dict, Keras layouts):
  embedding; encoder.* exactly as ForwardTransformer dense blocks (models.py docstring);
  prenet.d1.{w,b}, prenet.d2.{w,b}                                  (DecoderPrenet, layers.py:420-443)
  decoder.ln.{gamma,beta}, decoder.pos_scalar                       (CrossAttentionBlocks, layers.py:381-417)
  decoder.b{i}.sa.{wq,wk,wv,wo}.{w,b}, decoder.b{i}.sa.ln.{gamma,beta}   (SelfAttentionResNorm, layers.py:198-211)
  decoder.b{i}.ca.{wq,wk,wv,wo}.{w,b}, decoder.b{i}.ca.ln.{gamma,beta}   (CrossAttentionResnorm, layers.py:315-327)
  decoder.b{i}.ffn1.{w,b}, decoder.b{i}.ffn2.{w,b}, decoder.b{i}.ln2.{gamma,beta}  (FFNResNorm, layers.py:82-102)
  final_proj.{w,b}  (d, mel*max_r);  postnet.stop.{w,b} (mel,3);  postnet.mel.{w,b} (mel,mel)   (layers.py:446-460)
"""
from __future__ import annotations

from typing import Dict

import torch

from .. import lib
from .models import LN_EPS, ForwardTransformer, _on_device, _PackedLinear, _round_up
from .transformer_utils import mask_from_lengths, positional_encoding

ALIGNER_VOCAB = 129  # 126 symbols + pad + start + end (reference: data/text/tokenizer.py:17-26 with add_start_end=True)


class Aligner(ForwardTransformer):
    def __init__(self,
                 encoder_model_dimension: int,
                 decoder_model_dimension: int,
                 encoder_num_heads: list,
                 decoder_num_heads: list,
                 encoder_max_position_encoding: int,
                 decoder_max_position_encoding: int,
                 encoder_prenet_dimension: int,
                 decoder_prenet_dimension: int,
                 dropout_rate: float,
                 mel_start_value: float,
                 mel_end_value: float,
                 mel_channels: int,
                 phoneme_language: str = 'en-us',
                 with_stress: bool = True,
                 decoder_prenet_dropout: float = 0.1,
                 model_breathing: bool = False,
                 encoder_feed_forward_dimension: int = None,
                 decoder_feed_forward_dimension: int = None,
                 max_r: int = 10,
                 debug=False,
                 **kwargs):
        loc = dict(locals())
        self.config = {k: v for k, v in loc.items() if k not in ('self', 'kwargs', '__class__')}
        self.config.update(kwargs)
        if int(encoder_prenet_dimension) != int(encoder_model_dimension):
            raise ValueError('the embedding (encoder prenet) feeds the encoder blocks directly: dimensions must match '
                             '(reference: model/models.py:53-65)')
        self.mel_channels = int(mel_channels)
        self.vocab_size = int(kwargs.get('vocab_size', ALIGNER_VOCAB))
        self.device = torch.device(kwargs.get('device', 'cuda:0'))
        self.precision = kwargs.get('precision', 'bf16x3')
        self.impl = kwargs.get('impl', 'tcgen05')
        self.attention_precision = kwargs.get('attention_precision', 'fp16' if self.precision == 'bf16x3' else 'bf16')
        self.return_attention_weights = True      # attention maps are model outputs (models.py:150-153, 297)
        self._weights_all = True
        self.debug = debug
        self.alphabet = kwargs.get('alphabet')
        self.train_dropout = bool(kwargs.get('train_dropout', True))  # False: deterministic training step (parity tests)
        # replay the teacher-forced validation step / training step as CUDA graphs per input shape (the steps are ~90 / ~430
        # dependent launches of small kernels: host-launch bound when issued eagerly)
        self.cuda_graphs = bool(kwargs.get('cuda_graphs', False))
        self.train_graphs = bool(kwargs.get('train_graphs', False))
        self._val_graphs = {}
        self._graph_pool = None
        self.max_r = int(max_r)
        self.r = int(max_r)                        # models.py:46 -- starts at max_r, lowered by the schedule via set_constants
        self.stop_prob_index = 2
        self.force_encoder_diagonal = False
        self.force_decoder_diagonal = False
        self.stop_scaling = float(kwargs.get('stop_loss_scaling', 8.0))
        self.start_vec = torch.full((1, self.mel_channels), float(mel_start_value))
        self.end_vec = torch.full((1, self.mel_channels), float(mel_end_value))
        self._stacks = {
            'encoder': dict(d=int(encoder_model_dimension), heads=list(encoder_num_heads), n_dense=len(encoder_num_heads),
                            ffn=encoder_feed_forward_dimension, filters=[], kernel=None, max_pos=int(encoder_max_position_encoding)),
            'decoder': dict(d=int(decoder_model_dimension), heads=list(decoder_num_heads), n_dense=len(decoder_num_heads),
                            ffn=decoder_feed_forward_dimension, filters=[], kernel=None, max_pos=int(decoder_max_position_encoding)),
        }
        self.weights: Dict[str, torch.Tensor] = {}
        self._packed = None
        self._prof = None
        self.optimizer = None
        self.loss_weights = [1., 1.]
        self._engine = None
        self._drop_seed = 0
        self._step = 0
        self._init_weights(seed=int(kwargs.get('seed', 42)))

    # ------------------------------------------------------------------------------------------------
    def _param_shapes(self) -> Dict[str, tuple]:
        c = self.config
        enc, dec = self._stacks['encoder'], self._stacks['decoder']
        d_enc, d_dec, mel = enc['d'], dec['d'], self.mel_channels
        sh = {'embedding': (self.vocab_size, d_enc)}

        def mha(pre, d_q, d_kv, d):
            sh[pre + 'wq.w'], sh[pre + 'wq.b'] = (d_q, d), (d,)
            sh[pre + 'wk.w'], sh[pre + 'wk.b'] = (d_kv, d), (d,)
            sh[pre + 'wv.w'], sh[pre + 'wv.b'] = (d_kv, d), (d,)
            sh[pre + 'wo.w'], sh[pre + 'wo.b'] = (d_q + d, d), (d,)

        def ln(pre, n):
            sh[pre + '.gamma'], sh[pre + '.beta'] = (n,), (n,)

        def lin(pre, fin, fout):
            sh[pre + '.w'], sh[pre + '.b'] = (fin, fout), (fout,)

        ln('encoder.ln', d_enc)
        sh['encoder.pos_scalar'] = ()
        for i, _ in enumerate(enc['heads']):
            pre = f'encoder.b{i}.'
            mha(pre, d_enc, d_enc, d_enc)
            ln(pre + 'ln1', d_enc)
            lin(pre + 'ffn1', d_enc, int(enc['ffn']))
            lin(pre + 'ffn2', int(enc['ffn']), d_enc)
            ln(pre + 'ln2', d_enc)
        lin('prenet.d1', mel, int(c['decoder_prenet_dimension']))
        lin('prenet.d2', int(c['decoder_prenet_dimension']), d_dec)
        ln('decoder.ln', d_dec)
        sh['decoder.pos_scalar'] = ()
        for i, _ in enumerate(dec['heads']):
            pre = f'decoder.b{i}.'
            mha(pre + 'sa.', d_dec, d_dec, d_dec)
            ln(pre + 'sa.ln', d_dec)
            mha(pre + 'ca.', d_dec, d_enc, d_dec)
            ln(pre + 'ca.ln', d_dec)
            lin(pre + 'ffn1', d_dec, int(dec['ffn']))
            lin(pre + 'ffn2', int(dec['ffn']), d_dec)
            ln(pre + 'ln2', d_dec)
        lin('final_proj', d_dec, mel * self.max_r)
        lin('postnet.stop', mel, 3)
        lin('postnet.mel', mel, mel)
        return sh

    # ------------------------------------------------------------------------------------------------
    def _prepare(self):
        if self._packed is not None and self._packed['precision'] == self.precision:
            return self._packed
        lib.load()
        W, sp = self.weights, self._split
        P = {'precision': self.precision}
        enc, dec = self._stacks['encoder'], self._stacks['decoder']
        d_enc, d_dec, mel = enc['d'], dec['d'], self.mel_channels
        if d_enc % 64 or d_dec % 64 or int(enc['ffn']) % 64 or int(dec['ffn']) % 64 or int(self.config['decoder_prenet_dimension']) % 64:
            raise lib.TtsbError('Aligner: model / feed-forward / prenet dimensions must be multiples of 64 (GEMM K blocks)')
        P['encoder.pe'] = self._prepare_pe('encoder')
        for i, _ in enumerate(enc['heads']):
            pre = f'encoder.b{i}.'
            wqkv = torch.cat([W[pre + 'wq.w'], W[pre + 'wk.w'], W[pre + 'wv.w']], dim=1)
            bqkv = torch.cat([W[pre + 'wq.b'], W[pre + 'wk.b'], W[pre + 'wv.b']])
            P[pre + 'qkv'] = _PackedLinear(wqkv, bqkv, [d_enc], sp, block_n=d_enc if d_enc <= 256 else d_enc // 2)
            P[pre + 'wo'] = _PackedLinear(W[pre + 'wo.w'], W[pre + 'wo.b'], [d_enc, d_enc], sp, single_tile=True)
            P[pre + 'ffn1'] = _PackedLinear(W[pre + 'ffn1.w'], W[pre + 'ffn1.b'], [d_enc], sp)
            P[pre + 'ffn2'] = _PackedLinear(W[pre + 'ffn2.w'], W[pre + 'ffn2.b'], [int(enc['ffn'])], sp, single_tile=True)
        # K = mel_channels (80) is padded with zero rows to one 128-wide K block pair
        self._mel_k = _round_up(mel, 64)

        def pad_rows(w):
            out = torch.zeros((self._mel_k, w.shape[1]), dtype=w.dtype, device=w.device)
            out[:w.shape[0]] = w
            return out

        P['prenet.d1'] = _PackedLinear(pad_rows(W['prenet.d1.w']), W['prenet.d1.b'], [self._mel_k], sp)
        P['prenet.d2'] = _PackedLinear(W['prenet.d2.w'], W['prenet.d2.b'], [int(self.config['decoder_prenet_dimension'])], sp)
        for i, _ in enumerate(dec['heads']):
            pre = f'decoder.b{i}.'
            s = pre + 'sa.'
            wqkv = torch.cat([W[s + 'wq.w'], W[s + 'wk.w'], W[s + 'wv.w']], dim=1)
            bqkv = torch.cat([W[s + 'wq.b'], W[s + 'wk.b'], W[s + 'wv.b']])
            P[s + 'qkv'] = _PackedLinear(wqkv, bqkv, [d_dec], sp, block_n=d_dec if d_dec <= 256 else d_dec // 2)
            P[s + 'wo'] = _PackedLinear(W[s + 'wo.w'], W[s + 'wo.b'], [d_dec, d_dec], sp, single_tile=True)
            c = pre + 'ca.'
            P[c + 'q'] = _PackedLinear(W[c + 'wq.w'], W[c + 'wq.b'], [d_dec], sp)
            P[c + 'kv'] = _PackedLinear(torch.cat([W[c + 'wk.w'], W[c + 'wv.w']], dim=1), torch.cat([W[c + 'wk.b'], W[c + 'wv.b']]),
                                        [d_enc], sp, block_n=d_dec if d_dec <= 256 else d_dec // 2)
            P[c + 'wo'] = _PackedLinear(W[c + 'wo.w'], W[c + 'wo.b'], [d_dec, d_dec], sp, single_tile=True)
            P[pre + 'ffn1'] = _PackedLinear(W[pre + 'ffn1.w'], W[pre + 'ffn1.b'], [d_dec], sp)
            P[pre + 'ffn2'] = _PackedLinear(W[pre + 'ffn2.w'], W[pre + 'ffn2.b'], [int(dec['ffn'])], sp, single_tile=True)
        # Postnet: mel (80) and stop (3) heads share one GEMM over the padded linear frames
        w_post = pad_rows(torch.cat([W['postnet.mel.w'], W['postnet.stop.w']], dim=1))
        P['postnet'] = _PackedLinear(w_post, torch.cat([W['postnet.mel.b'], W['postnet.stop.b']]), [self._mel_k], sp)
        self._packed = P
        self._final_proj = {}
        self._pe_r = {}
        return P

    def _final_proj_r(self, r: int) -> _PackedLinear:
        """Dense(mel*max_r) followed by [:, :, :r*mel] (models.py:146): only the first r*mel output columns are computed."""
        if r not in self._final_proj:
            n = r * self.mel_channels
            self._final_proj[r] = _PackedLinear(self.weights['final_proj.w'][:, :n].contiguous(), self.weights['final_proj.b'][:n].contiguous(),
                                                [self._stacks['decoder']['d']], self._split)
        return self._final_proj[r]

    def _decoder_pe(self, r: int) -> torch.Tensor:
        """pos_encoding[:, :T*r:r] (layers.py:409) as a dense table so row t of the table is position t*r."""
        if not hasattr(self, '_pe_r'):
            self._pe_r = {}
        if r not in self._pe_r:
            st = self._stacks['decoder']
            self._pe_r[r] = positional_encoding(st['max_pos'], st['d'])[0][::r].to(self.device).contiguous()
        return self._pe_r[r]

    # ------------------------------------------------------------------------------------------------
    def _mha(self, B, T, H, dh, q_buf, ld_q, q_col0, kv_buf, ld_kv, Tk, k_col0, v_col0, lens, causal, weights):
        d = H * dh
        _, at_hi, at_lo = self._act(B, T, d, f32=False)
        ap = self.attention_precision
        if ap == 'bf16x3':
            raise lib.TtsbError("Aligner attention runs in the single-pass modes ('fp16' / 'bf16'): head dim 256 needs them")
        m = lib.MhaArgs()
        m.B, m.T, m.H, m.dh = B, T, H, dh
        m.qk_hi = q_buf.data_ptr()
        m.ld_qk, m.q_col0, m.k_col0, m.v_col0 = ld_q, q_col0, k_col0, v_col0
        if kv_buf is not None:
            m.kv_hi = kv_buf.data_ptr()
            m.ld_kv, m.Tk = ld_kv, Tk
        m.kv_len = lens.data_ptr()
        m.out_hi = at_hi.data_ptr()
        m.out_lo = at_lo.data_ptr() if at_lo is not None else None
        m.ld_out = d
        m.causal = int(causal)
        m.full_queries = 1
        wts = None
        if weights:
            wts = torch.empty((B, H, T, Tk if kv_buf is not None else T), dtype=torch.float32, device=self.device)
            m.weights_out = wts.data_ptr()
            m.weights_all = 1
        m.precision = {'fp16': lib.PREC_FP16, 'bf16': lib.PREC_BF16}[ap]
        m.impl = self._impl
        lib.mha_fwd(m)
        return (at_hi, at_lo), wts

    def _cadb(self, P, i: int, x, enc, enc_len, dec_len, B: int, T: int, Tp: int):
        """CrossAttentionDenseBlock (layers.py:330-349): no row masks inside the block."""
        W = self.weights
        dec, d_enc = self._stacks['decoder'], self._stacks['encoder']['d']
        d, H = dec['d'], dec['heads'][i]
        dh = d // H
        pre = f'decoder.b{i}.'
        f16 = self.attention_precision == 'fp16'
        adt = torch.float16 if f16 else torch.bfloat16
        x_f, x_hi, x_lo = x
        # ---- masked (look-ahead + padding) self-attention, residual, LayerNorm
        qkv = P[pre + 'sa.qkv']
        qk = torch.empty((B, T, qkv.n_pad), dtype=adt, device=self.device)
        self._gemm(qkv, B, T, [(x_hi, x_lo, d, 0)], [0], [0], out_hi=qk, out_fp16=f16)
        (a_hi, a_lo), _ = self._mha(B, T, H, dh, qk, qkv.n_pad, 0, None, 0, T, d, 2 * d, dec_len, True, False)
        y = self._act(B, T, d)
        self._gemm(P[pre + 'sa.wo'], B, T, [(x_hi, x_lo, d, 0), (a_hi, a_lo, d, 0)], [0, 1], [0, 0], residual=x_f,
                   ln=(W[pre + 'sa.ln.gamma'], W[pre + 'sa.ln.beta']), out_f32=y[0], out_hi=y[1], out_lo=y[2])
        # ---- cross-attention onto the encoder output (keys masked by the encoder padding mask), residual, LayerNorm
        pq, pkv = P[pre + 'ca.q'], P[pre + 'ca.kv']
        qb = torch.empty((B, T, pq.n_pad), dtype=adt, device=self.device)
        kvb = torch.empty((B, Tp, pkv.n_pad), dtype=adt, device=self.device)
        self._gemm(pq, B, T, [(y[1], y[2], d, 0)], [0], [0], out_hi=qb, out_fp16=f16)
        self._gemm(pkv, B, Tp, [(enc[1], enc[2], d_enc, 0)], [0], [0], out_hi=kvb, out_fp16=f16)
        (c_hi, c_lo), wts = self._mha(B, T, H, dh, qb, pq.n_pad, 0, kvb, pkv.n_pad, Tp, 0, d, enc_len, False, True)
        z = self._act(B, T, d)
        self._gemm(P[pre + 'ca.wo'], B, T, [(y[1], y[2], d, 0), (c_hi, c_lo, d, 0)], [0, 1], [0, 0], residual=y[0],
                   ln=(W[pre + 'ca.ln.gamma'], W[pre + 'ca.ln.beta']), out_f32=z[0], out_hi=z[1], out_lo=z[2])
        # ---- feed-forward, residual, LayerNorm
        f1 = P[pre + 'ffn1']
        _, h_hi, h_lo = self._act(B, T, f1.n_pad, f32=False)
        self._gemm(f1, B, T, [(z[1], z[2], d, 0)], [0], [0], relu=True, out_hi=h_hi, out_lo=h_lo)
        o = self._act(B, T, d)
        self._gemm(P[pre + 'ffn2'], B, T, [(h_hi, h_lo, f1.n_pad, 0)], [0], [0], residual=z[0],
                   ln=(W[pre + 'ln2.gamma'], W[pre + 'ln2.beta']), out_f32=o[0], out_hi=o[1], out_lo=o[2])
        return o, wts

    # ------------------------------------------------------------------------------------------------
    # reference API
    # ------------------------------------------------------------------------------------------------
    def _call_encoder(self, inputs, training=False):
        """models.py:127-133 -> (encoder output triple, padding mask, attention weights, lengths)."""
        if training:
            raise lib.TtsbError('Aligner.call runs the inference path (training=False); dropout + backward live in train_step (aligner_training.AlignerTrainEngine)')
        P, W, dev = self._prepare(), self.weights, self.device
        x = torch.as_tensor(inputs).to(device=dev, dtype=torch.int32).contiguous()
        if x.dim() != 2:
            raise ValueError('input tokens must have shape (batch, length)')
        B, Tp = x.shape
        d = self._stacks['encoder']['d']
        enc_len = torch.empty((B,), dtype=torch.int32, device=dev)
        lib.phoneme_lengths(x, 0, enc_len)
        h = self._act(B, Tp, d)
        lib.embed_ln_pe_fwd(x, W['embedding'], W['encoder.ln.gamma'], W['encoder.ln.beta'], P['encoder.pe'],
                            W['encoder.pos_scalar'].reshape(1), LN_EPS, h[0], h[1], h[2])
        attn = {}
        for i in range(len(self._stacks['encoder']['heads'])):
            h = self._block(P, 'encoder', i, h, enc_len, B, Tp, attn, f'Encoder_DenseBlock{i + 1}_SelfAttention')
        return h, mask_from_lengths(enc_len, Tp), attn, enc_len

    def _call_decoder(self, encoder_output, targets, encoder_padding_mask, training=False, enc_len=None):
        """models.py:135-154.  encoder_output: the activation triple returned by _call_encoder."""
        if training:
            raise lib.TtsbError('Aligner.call runs the inference path (training=False); dropout + backward live in train_step (aligner_training.AlignerTrainEngine)')
        P, W, dev = self._prepare(), self.weights, self.device
        tgt = torch.as_tensor(targets).to(device=dev, dtype=torch.float32).contiguous()
        B, T, mel = tgt.shape
        if mel != self.mel_channels:
            raise ValueError(f'targets must have {self.mel_channels} channels')
        r = int(self.r)
        dec = self._stacks['decoder']
        d = dec['d']
        Tp = encoder_output[0].shape[1]
        if enc_len is None:
            enc_len = (1.0 - encoder_padding_mask[:, 0, 0, :]).sum(dim=1).to(torch.int32).contiguous()
        if T * r > dec['max_pos']:
            raise ValueError('target length * r exceeds decoder_max_position_encoding')
        # value-derived mel padding mask (transformer_utils.py:29-32) as per-row lengths (batches are padded at the end)
        dec_len = torch.empty((B,), dtype=torch.int32, device=dev)
        lib.mel_lengths(tgt, 0.0, dec_len)
        # ---- DecoderPrenet (layers.py:420-443): relu Dense -> relu Dense
        k = self._mel_k
        padded = torch.zeros((B, T, k), dtype=torch.float32, device=dev)
        padded[..., :mel] = tgt
        t_hi, t_lo = lib.split_bf16(padded, self._split)
        p1 = P['prenet.d1']
        _, h_hi, h_lo = self._act(B, T, p1.n_pad, f32=False)
        self._gemm(p1, B, T, [(t_hi, t_lo, k, 0)], [0], [0], relu=True, out_hi=h_hi, out_lo=h_lo)
        pre_out = torch.empty((B, T, d), dtype=torch.float32, device=dev)
        self._gemm(P['prenet.d2'], B, T, [(h_hi, h_lo, p1.n_pad, 0)], [0], [0], relu=True, out_f32=pre_out)
        # ---- CrossAttentionBlocks prologue (layers.py:406-410): LN(inputs) + scalar * PE[:, :T*r:r]
        idx = torch.arange(T, dtype=torch.int32, device=dev)[None, :].expand(B, T).contiguous()
        x = self._act(B, T, d)
        lib.expand_ln_pe_fwd(pre_out, idx, W['decoder.ln.gamma'], W['decoder.ln.beta'], self._decoder_pe(r),
                             W['decoder.pos_scalar'].reshape(1), LN_EPS, x[0], x[1], x[2])
        attn = {}
        n = len(dec['heads'])
        for i in range(n):
            x, wts = self._cadb(P, i, x, encoder_output, enc_len, dec_len, B, T, Tp)
            key = 'Decoder_LastBlock_CrossAttention' if i == n - 1 else f'Decoder_DenseBlock{i + 1}_CrossAttention'
            attn[key] = wts
        # ---- FinalProj[:, :, :r*mel] -> (B, T*r, mel) -> Postnet (models.py:146-150)
        fp = self._final_proj_r(r)
        lin = torch.empty((B, T, r * mel), dtype=torch.float32, device=dev)
        self._gemm(fp, B, T, [(x[1], x[2], d, 0)], [0], [0], out_f32=lin, ld_out=r * mel)
        linear = lin.view(B, T * r, mel)
        lpad = torch.zeros((B, T * r, k), dtype=torch.float32, device=dev)
        lpad[..., :mel] = linear
        l_hi, l_lo = lib.split_bf16(lpad, self._split)
        pn = P['postnet']
        post = torch.empty((B, T * r, pn.n_pad), dtype=torch.float32, device=dev)
        self._gemm(pn, B, T * r, [(l_hi, l_lo, k, 0)], [0], [0], out_f32=post)
        return {'mel': post[..., :mel].contiguous(), 'stop_prob': post[..., mel:mel + 3].contiguous(),
                'decoder_attention': attn, 'decoder_output': x[0], 'linear': linear,
                'mel_mask': mask_from_lengths(dec_len, T), 'mel_lengths': dec_len}

    @_on_device
    def call(self, inputs, targets, training=False):
        """models.py:294-298."""
        enc, padding_mask, enc_attn, enc_len = self._call_encoder(inputs, training)
        out = self._call_decoder(enc, targets, padding_mask, training, enc_len=enc_len)
        out.update({'encoder_attention': enc_attn, 'text_mask': padding_mask, 'text_lengths': enc_len,
                    'encoder_output': enc[0]})
        return out

    __call__ = call

    def _forward(self, inp, output):
        return self.call(inp, output, training=False)

    @_on_device
    def _forward_encoder(self, inputs):
        enc, mask, attn, _ = self._call_encoder(inputs, training=False)
        return enc, mask, attn

    @_on_device
    def _forward_decoder(self, encoder_output, targets, encoder_padding_mask):
        return self._call_decoder(encoder_output, targets, encoder_padding_mask, training=False)

    @_on_device
    def _gta_forward(self, inp, tar, stop_prob, training=False):
        """models.py:168-210 (forward + losses).  Returns (model_out, None): there is no tape here."""
        tar = torch.as_tensor(tar).to(device=self.device, dtype=torch.float32)
        stop = torch.as_tensor(stop_prob).to(device=self.device, dtype=torch.int32)
        tar_inp, tar_real, tar_stop = tar[:, :-1], tar[:, 1:].contiguous(), stop[:, 1:].contiguous()
        mel_len = tar_inp.shape[1]
        tar_mel = tar_inp[:, 0::self.r, :].contiguous()
        out = self.call(inp, tar_mel, training=training)
        dev = self.device
        B = tar.shape[0]
        l_mel = torch.zeros((1,), dtype=torch.float32, device=dev)
        l_stop = torch.zeros((1,), dtype=torch.float32, device=dev)
        lib.mae_loss(out['mel'], B, out['mel'].shape[1], mel_len, self.mel_channels, tar_real, 1.0, l_mel, None)
        lib.scaled_ce_loss(out['stop_prob'], mel_len, 3, tar_stop, self.stop_prob_index, self.stop_scaling, l_stop)
        d_loss = torch.zeros((1,), dtype=torch.float32, device=dev)
        norm = 1.0
        if self.force_decoder_diagonal:
            for w in out['decoder_attention'].values():
                lib.diag_loss(w, out['mel_lengths'], out['text_lengths'], d_loss)
            norm += len(out['decoder_attention'])
        if self.force_encoder_diagonal:
            for w in out['encoder_attention'].values():
                lib.diag_loss(w, out['text_lengths'], out['text_lengths'], d_loss)
            norm += len(out['encoder_attention'])
        d_loss = d_loss / norm
        loss = self.loss_weights[0] * l_mel + self.loss_weights[1] * l_stop + d_loss
        out.update({'loss': loss[0], 'losses': {'mel': l_mel[0], 'stop_prob': l_stop[0], 'diag_loss': d_loss[0]}})
        return out, None

    def _val_step(self, inp, tar, stop_prob):
        if self.cuda_graphs:
            return self._val_step_graphed(inp, tar, stop_prob)
        return self._gta_forward(inp, tar, stop_prob, training=False)[0]

    @_on_device
    def _val_step_graphed(self, inp, tar, stop_prob):
        """The validation step captured once per (shapes, r, diagonal flags) and replayed; outputs are copied out of the
        graph's static buffers."""
        inp, tar, stop_prob = torch.as_tensor(inp), torch.as_tensor(tar), torch.as_tensor(stop_prob)
        self._prepare()
        key = (tuple(inp.shape), tuple(tar.shape), self.r, self.force_encoder_diagonal, self.force_decoder_diagonal, id(self._packed))
        ent = self._val_graphs.get(key)
        if ent is None:
            dev = self.device
            ins = [inp.to(device=dev, dtype=torch.int32).contiguous().clone(), tar.to(device=dev, dtype=torch.float32).contiguous().clone(),
                   stop_prob.to(device=dev, dtype=torch.int32).contiguous().clone()]
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._gta_forward(*ins, training=False)
            torch.cuda.current_stream().wait_stream(side)
            if self._graph_pool is None:
                self._graph_pool = torch.cuda.graph_pool_handle()
            g = torch.cuda.CUDAGraph()
            n0 = lib.launch_count()
            with torch.cuda.graph(g, pool=self._graph_pool):
                out = self._gta_forward(*ins, training=False)[0]
            if len(self._val_graphs) >= 4:
                self._val_graphs.pop(next(iter(self._val_graphs)))
            ent = self._val_graphs[key] = {'ins': ins, 'g': g, 'out': out, 'n': lib.launch_count() - n0}
        else:
            for dst, src in zip(ent['ins'], (inp, tar, stop_prob)):
                dst.copy_(src, non_blocking=True)
        ent['g'].replay()
        lib.add_launch_count(ent['n'])

        def cp(v):
            if torch.is_tensor(v):
                return v.clone()
            if isinstance(v, dict):
                return {k: cp(x) for k, x in v.items()}
            return v
        return cp(ent['out'])

    val_step = _val_step

    def _get_engine(self):
        if self._engine is None:
            from .aligner_training import AlignerTrainEngine
            self._engine = AlignerTrainEngine(self)
        return self._engine

    @_on_device
    def _train_step(self, inp, tar, stop_prob, data_parallel: bool = False):
        """models.py:212-216: teacher-forced forward (dropout on, single-pass bf16), hand-written backward, Adam.
        The returned dictionary has the losses and outputs; attention maps are not materialised in fp32 on this path."""
        if self.optimizer is None:
            self._compile(self.stop_scaling)
        eng = self._get_engine()
        sync = None
        if data_parallel:
            from ..utils.data_parallel import make_grad_sync
            sync = make_grad_sync(eng.flat_g)
        if self.train_graphs and sync is None:
            out = eng.step_graphed(inp, tar, stop_prob)
        else:
            out = eng.forward_backward(inp, tar, stop_prob, training=True, sync=sync)
        scale = sync.finish() if sync is not None else 1.0
        eng.apply_adam(self.optimizer, grad_scale=scale)
        return out

    train_step = _train_step

    def encode_text(self, text):
        """models.py:338-340: text -> token ids through the attached text pipeline (the espeak phonemizer is external)."""
        tp = getattr(self, 'text_pipeline', None)
        if tp is None:
            raise NotImplementedError('text encoding needs the espeak phonemizer, which is outside the built path; '
                                      'pass token ids with encode=False or attach a text_pipeline')
        return tp(text)

    def predict(self, inp, max_length=1000, encode=True, verbose=True):
        """models.py:271-292: autoregressive decoding of one token row.  As in the reference the encoder runs once and the
        decoder is re-run on the whole prefix every iteration; the prefix grows by the last predicted frame, the returned mel
        by the last r frames, and decoding stops when the arg-max of the last stop distribution is `stop_prob_index`.
        One host read per iteration (the stop decision), as `int(tf.argmax(...))` is in the reference."""
        if encode:
            inp = self.encode_text(inp)
        return self._predict_tokens(inp, max_length, verbose)

    @_on_device
    def _predict_tokens(self, inp, max_length, verbose):
        dev = self.device
        inp = torch.as_tensor(inp).to(device=dev, dtype=torch.int32).reshape(1, -1)
        output = self.start_vec.to(device=dev, dtype=torch.float32).reshape(1, 1, self.mel_channels)
        output_concat = output.clone()
        out_dict = {}
        enc, padding_mask, enc_attn, enc_len = self._call_encoder(inp, training=False)
        r = int(self.r)
        for _ in range(int(max_length // r) + 1):
            mo = self._call_decoder(enc, output, padding_mask, training=False, enc_len=enc_len)
            output = torch.cat([output, mo['mel'][:1, -1:, :]], dim=-2)
            output_concat = torch.cat([output_concat, mo['mel'][:1, -r:, :]], dim=-2)
            out_dict = {'mel': output_concat[0, 1:, :], 'decoder_attention': mo['decoder_attention'], 'encoder_attention': enc_attn}
            if int(torch.argmax(mo['stop_prob'][:, -1], dim=-1)) == self.stop_prob_index:
                if verbose:
                    print('Stopping')
                break
        return out_dict

    def _compile(self, stop_scaling=8.0, optimizer=None):
        """models.py:222-227."""
        from .training import Adam
        self.loss_weights = [1., 1.]
        self.stop_scaling = float(stop_scaling)
        self.optimizer = optimizer if optimizer is not None else Adam(1.0e-4)

    def _set_r(self, r):
        self.r = int(r)

    def set_constants(self, learning_rate: float = None, reduction_factor: float = None, decoder_prenet_dropout: float = None,
                      force_encoder_diagonal: bool = None, force_decoder_diagonal: bool = None):
        """models.py:300-312."""
        if reduction_factor is not None:
            self._set_r(reduction_factor)
        if force_encoder_diagonal is not None:
            self.force_encoder_diagonal = bool(force_encoder_diagonal)
        if force_decoder_diagonal is not None:
            self.force_decoder_diagonal = bool(force_decoder_diagonal)

    @property
    def step(self) -> int:
        return int(self.optimizer.iterations) if self.optimizer is not None else 0

    @classmethod
    def from_config(cls, config: dict, max_r: int = 10):
        """models.py:320-341."""
        keys = ('encoder_model_dimension', 'decoder_model_dimension', 'encoder_num_heads', 'decoder_num_heads',
                'encoder_max_position_encoding', 'decoder_max_position_encoding', 'encoder_prenet_dimension',
                'decoder_prenet_dimension', 'dropout_rate', 'mel_start_value', 'mel_end_value', 'mel_channels',
                'phoneme_language', 'with_stress', 'decoder_prenet_dropout', 'model_breathing',
                'encoder_feed_forward_dimension', 'decoder_feed_forward_dimension')
        kw = {k: config[k] for k in keys if k in config}
        extra = {k: config[k] for k in ('vocab_size', 'precision', 'attention_precision', 'impl', 'device', 'seed', 'stop_loss_scaling', 'train_dropout') if k in config}
        return cls(max_r=int(config.get('max_r', max_r)), debug=config.get('debug', False), **kw, **extra)
