"""Training step of the Aligner (reference: model/models.py:168-216 ``_gta_forward`` + ``_train_step``): teacher-forced
forward in single-pass bf16 with saved activations, hand-written backward, Keras-form Adam -- on the same kernels as the
ForwardTransformer engine (training.py).  The encoder blocks reuse TrainEngine._block_fwd/_block_bwd unchanged; this file
adds the CrossAttentionDenseBlock (layers.py:330-349), DecoderPrenet, FinalProj / Postnet and the Aligner losses.

Differences to the ForwardTransformer blocks that matter for the backward pass:
  * decoder rows are never re-masked (no ``row_len`` in the LayerNorm GEMMs / LayerNorm backward);
  * decoder self-attention uses the look-ahead + padding mask and keeps every query row (softmax flags);
  * cross-attention reads K|V from a second GEMM over the encoder output; its gradient flows back into the encoder
    output from every decoder block (accumulated through the residual input of the data-gradient GEMM);
  * the diagonal loss (models.py:186-207) is taken on the post-dropout attention probabilities and adds a term to dP.
"""
from __future__ import annotations

import math

import torch

from .. import lib
from .models import LN_EPS, _packed_empty, _round_up
from .training import TrainEngine
from .transformer_utils import mask_from_lengths

FULLQ = lib.SOFTMAX_FULL_QUERIES
CAUSAL = lib.SOFTMAX_CAUSAL


class AlignerTrainEngine(TrainEngine):
    # ------------------------------------------------------------------------------------------------
    # packed operands
    # ------------------------------------------------------------------------------------------------
    def _build_packs(self):
        m = self.model
        W = m.weights
        P, descs = {}, []
        dev = self.dev

        def desc(src, dst_ptr, R, R_pad, C_cols, cb, cb_valid, sr, s_outer, s_inner, dst_ld, f32=0):
            d_ = lib.PackDesc()
            d_.src, d_.dst = src.data_ptr() if torch.is_tensor(src) else src, dst_ptr
            d_.R, d_.R_pad, d_.C_cols, d_.cb, d_.cb_valid = R, R_pad, C_cols, cb, cb_valid
            d_.sr, d_.s_outer, d_.s_inner, d_.dst_ld, d_.dst_f32 = sr, s_outer, s_inner, dst_ld, f32
            descs.append(d_)

        def fwd(key, parts, K, seg_k, single=False, block_n=None, k_valid=None):
            """parts: [(w (Kv,Ni) view, b (Ni))] concatenated along N -- forward packing [N_pad, K]; rows k_valid..K of the
            contraction are zero (the 80 mel channels are fed as a 128-wide K block)."""
            kv = K if k_valid is None else k_valid
            N = sum(w.shape[-1] for w, _ in parts)
            pl = _packed_empty(K, N, seg_k, dev, single_tile=single, block_n=block_n)
            if kv != K:
                pl.w_hi.zero_()
            row = 0
            for i, (w, b) in enumerate(parts):
                Ni = w.shape[-1]
                last = i == len(parts) - 1
                desc(w, pl.w_hi.data_ptr() + 2 * row * K, Ni, (pl.n_pad - row) if last else Ni, K, K, kv, 1, 0, Ni, K)
                desc(b, pl.bias.data_ptr() + 4 * row, 1, 1, (pl.n_pad - row) if last else Ni, pl.n_pad, Ni, 0, 0, 1, pl.n_pad, f32=1)
                row += Ni
            P[key] = pl

        def dgrad_dense(key, parts, K):
            """parts: [w (K,Ni)] concatenated along N -- data-gradient packing [K_pad, N_pad] (contraction over N)."""
            N = sum(w.shape[-1] for w in parts)
            npad = _round_up(N, 64)
            pl = _packed_empty(npad, K, [npad], dev, bias=False)
            col = 0
            for i, w in enumerate(parts):
                Ni = w.shape[-1]
                last = i == len(parts) - 1
                width = (npad - col) if last else Ni
                desc(w, pl.w_hi.data_ptr() + 2 * col, K, pl.n_pad, width, width, Ni, Ni, 0, 1, npad)
                col += Ni
            P[key] = pl

        enc, dec = m._stacks['encoder'], m._stacks['decoder']
        d_enc, d_dec, mel = enc['d'], dec['d'], m.mel_channels
        for i, _ in enumerate(enc['heads']):
            pre = f'encoder.b{i}.'
            d = d_enc
            qkv = [(W[pre + n + '.w'], W[pre + n + '.b']) for n in ('wq', 'wk', 'wv')]
            fwd(pre + 'qkv', qkv, d, [d], block_n=d if d <= 256 else d // 2)
            dgrad_dense(pre + 'qkv.d', [w for w, _ in qkv], d)
            fwd(pre + 'wo', [(W[pre + 'wo.w'], W[pre + 'wo.b'])], 2 * d, [d, d], single=True)
            dgrad_dense(pre + 'wo.dx', [W[pre + 'wo.w'][:d]], d)
            dgrad_dense(pre + 'wo.da', [W[pre + 'wo.w'][d:]], d)
            F = int(enc['ffn'])
            fwd(pre + 'ffn1', [(W[pre + 'ffn1.w'], W[pre + 'ffn1.b'])], d, [d])
            fwd(pre + 'ffn2', [(W[pre + 'ffn2.w'], W[pre + 'ffn2.b'])], F, [F], single=True)
            dgrad_dense(pre + 'ffn1.d', [W[pre + 'ffn1.w']], d)
            dgrad_dense(pre + 'ffn2.d', [W[pre + 'ffn2.w']], F)
        kmel = _round_up(mel, 64)
        pd = int(m.config['decoder_prenet_dimension'])
        fwd('prenet.d1', [(W['prenet.d1.w'], W['prenet.d1.b'])], kmel, [kmel], k_valid=mel)
        fwd('prenet.d2', [(W['prenet.d2.w'], W['prenet.d2.b'])], pd, [pd])
        dgrad_dense('prenet.d2.d', [W['prenet.d2.w']], pd)
        d = d_dec
        for i, _ in enumerate(dec['heads']):
            pre = f'decoder.b{i}.'
            s = pre + 'sa.'
            qkv = [(W[s + n + '.w'], W[s + n + '.b']) for n in ('wq', 'wk', 'wv')]
            fwd(s + 'qkv', qkv, d, [d], block_n=d if d <= 256 else d // 2)
            dgrad_dense(s + 'qkv.d', [w for w, _ in qkv], d)
            fwd(s + 'wo', [(W[s + 'wo.w'], W[s + 'wo.b'])], 2 * d, [d, d], single=True)
            dgrad_dense(s + 'wo.dx', [W[s + 'wo.w'][:d]], d)
            dgrad_dense(s + 'wo.da', [W[s + 'wo.w'][d:]], d)
            c = pre + 'ca.'
            fwd(c + 'q', [(W[c + 'wq.w'], W[c + 'wq.b'])], d, [d])
            dgrad_dense(c + 'q.d', [W[c + 'wq.w']], d)
            kv = [(W[c + 'wk.w'], W[c + 'wk.b']), (W[c + 'wv.w'], W[c + 'wv.b'])]
            fwd(c + 'kv', kv, d_enc, [d_enc], block_n=d if d <= 256 else d // 2)
            dgrad_dense(c + 'kv.d', [w for w, _ in kv], d_enc)
            fwd(c + 'wo', [(W[c + 'wo.w'], W[c + 'wo.b'])], 2 * d, [d, d], single=True)
            dgrad_dense(c + 'wo.dx', [W[c + 'wo.w'][:d]], d)
            dgrad_dense(c + 'wo.da', [W[c + 'wo.w'][d:]], d)
            F = int(dec['ffn'])
            fwd(pre + 'ffn1', [(W[pre + 'ffn1.w'], W[pre + 'ffn1.b'])], d, [d])
            fwd(pre + 'ffn2', [(W[pre + 'ffn2.w'], W[pre + 'ffn2.b'])], F, [F], single=True)
            dgrad_dense(pre + 'ffn1.d', [W[pre + 'ffn1.w']], d)
            dgrad_dense(pre + 'ffn2.d', [W[pre + 'ffn2.w']], F)
        # FinalProj[:, :, :r*mel] per reduction factor (models.py:146); Postnet: mel | stop heads in one GEMM
        self._fp = {}
        for r in range(1, m.max_r + 1):
            n = r * mel
            w, b = W['final_proj.w'][:, :n], W['final_proj.b'][:n]
            N = n
            pl = _packed_empty(d, N, [d], dev)
            desc(w, pl.w_hi.data_ptr(), N, pl.n_pad, d, d, d, 1, 0, W['final_proj.w'].shape[1], d)
            desc(b, pl.bias.data_ptr(), 1, 1, pl.n_pad, pl.n_pad, N, 0, 0, 1, pl.n_pad, f32=1)
            npad = _round_up(N, 64)
            pd_ = _packed_empty(npad, d, [npad], dev, bias=False)
            desc(w, pd_.w_hi.data_ptr(), d, pd_.n_pad, npad, npad, N, W['final_proj.w'].shape[1], 0, 1, npad)
            self._fp[r] = (pl, pd_)
        post = [(W['postnet.mel.w'], W['postnet.mel.b']), (W['postnet.stop.w'], W['postnet.stop.b'])]
        fwd('postnet', post, kmel, [kmel], k_valid=mel)
        dgrad_dense('postnet.d', [w for w, _ in post], mel)
        P['encoder.pe'] = m._prepare_pe('encoder')
        self.P = P
        self._n_descs = len(descs)
        self._descs_dev = lib.upload_pack_descs(descs, dev)

    # ------------------------------------------------------------------------------------------------
    # generic attention core on materialised probabilities (self: q,k,v in one buffer; cross: q buffer + kv buffer)
    # ------------------------------------------------------------------------------------------------
    def _attn_fwd(self, B, H, dh, T, Tk, qb, q_ld, q_col, kb, k_ld, k_col, v_col, lens, flags):
        d = H * dh
        ldp = _round_up(Tk, 16)
        Z = B * H
        S = self._f32(Z, T, ldp)
        self._bgemm(B, H, T, Tk, dh, qb, (d, T, B), (q_ld, q_ld * T), (dh, 0, 0, q_col), kb, (d, Tk, B), (k_ld, k_ld * Tk),
                    (dh, 0, 0, k_col), alpha=1.0 / math.sqrt(dh), out_f32=S, ld_out=ldp, out_batch_stride=T * ldp, out_cols=ldp)
        P_pre = self._bf(Z, T, ldp)
        rate = self.drop_rate
        P_drop = self._bf(Z, T, ldp) if rate > 0 else P_pre
        site = self._site()
        lib.softmax_fwd(S, B, H, T, Tk, ldp, lens, rate, self.seed, site, P_pre, P_drop, flags=flags)
        del S
        out = self._bf(B, T, d)
        self._bgemm(B, H, T, dh, Tk, P_drop, (Tk, T, Z), (ldp, T * ldp), (0, 0, 1, 0), kb, (d, Tk, B), (k_ld, k_ld * Tk),
                    (dh, 0, 0, v_col, 1), out_bf16=out, ld_out=d, out_batch_stride=T * d, out_h_col=dh, out_by_b=1, out_cols=dh)
        return out, dict(P_pre=P_pre, P_drop=P_drop, ldp=ldp, site=site, flags=flags, out=out)

    def _attn_bwd(self, c, B, H, dh, T, Tk, dout, qb, q_ld, q_col, kb, k_ld, k_col, v_col, lens, dq_buf, dq_ld, dq_col, dkv_buf,
                  dkv_ld, dk_col, dv_col, diag=None):
        """dout: bf16 (B,T,d) gradient of the attention output.  Writes dQ into dq_buf[:, :, dq_col:], dK / dV into
        dkv_buf[:, :, dk_col:] / [dv_col:] (bf16).  diag = (grad_scale, q_len, k_len) adds the diagonal-loss term to dP."""
        d = H * dh
        ldp = c['ldp']
        Z = B * H
        dS = self._bf(Z, T, ldp)
        scale = 1.0 / math.sqrt(dh)
        if diag is not None:
            dP = self._f32(Z, T, ldp)
            self._bgemm(B, H, T, Tk, dh, dout, (d, T, B), (d, d * T), (dh, 0, 0, 0), kb, (d, Tk, B), (k_ld, k_ld * Tk), (dh, 0, 0, v_col),
                        out_f32=dP, ld_out=ldp, out_batch_stride=T * ldp, out_cols=ldp)
            lib.diag_loss_train(c['P_drop'], B, H, T, Tk, ldp, diag[1], diag[2], 0.0, self._scratch1, diag[0], dP)
            lib.softmax_bwd(c['P_pre'], dP, B, H, T, Tk, ldp, lens, scale, self.drop_rate, self.seed, c['site'], dS, flags=c['flags'])
            del dP
        else:  # dS out of the dP product's epilogue (D = dO . O), see TrainEngine._block_bwd
            D = self._f32(Z * T)
            lib.rowdot_heads(dout, c['out'], H, dh, D)
            self._bgemm(B, H, T, Tk, dh, dout, (d, T, B), (d, d * T), (dh, 0, 0, 0), kb, (d, Tk, B), (k_ld, k_ld * Tk), (dh, 0, 0, v_col),
                        out_bf16=dS, ld_out=ldp, out_batch_stride=T * ldp, out_cols=ldp,
                        softmax_bwd=(c['P_pre'], D, scale, self.drop_rate, self.seed, c['site'], c['flags'], lens, c['P_drop']))
        # dQ = dS K : A = dS (K-major over keys), B = K read MN-major
        self._bgemm(B, H, T, dh, Tk, dS, (Tk, T, Z), (ldp, T * ldp), (0, 0, 1, 0), kb, (d, Tk, B), (k_ld, k_ld * Tk),
                    (dh, 0, 0, k_col, 1), out_bf16=dq_buf, ld_out=dq_ld, out_batch_stride=T * dq_ld, out_h_col=dh, out_by_b=1,
                    out_cols=dh, out_ptr_off=dq_col)
        # dK = dS^T Q : A = dS read MN-major (= dS^T), B = Q read MN-major
        self._bgemm(B, H, Tk, dh, T, dS, (Tk, T, Z), (ldp, T * ldp), (0, 0, 1, 0, 1), qb, (d, T, B), (q_ld, q_ld * T),
                    (dh, 0, 0, q_col, 1), out_bf16=dkv_buf, ld_out=dkv_ld, out_batch_stride=Tk * dkv_ld, out_h_col=dh, out_by_b=1,
                    out_cols=dh, out_ptr_off=dk_col)
        # dV = P^T dO : A = P_drop read MN-major, B = dO read MN-major
        self._bgemm(B, H, Tk, dh, T, c['P_drop'], (Tk, T, Z), (ldp, T * ldp), (0, 0, 1, 0, 1), dout, (d, T, B), (d, d * T),
                    (dh, 0, 0, 0, 1), out_bf16=dkv_buf, ld_out=dkv_ld, out_batch_stride=Tk * dkv_ld, out_h_col=dh, out_by_b=1,
                    out_cols=dh, out_ptr_off=dv_col)

    # ------------------------------------------------------------------------------------------------
    # CrossAttentionDenseBlock
    # ------------------------------------------------------------------------------------------------
    def _cadb_fwd(self, i, x_f, x_bf, enc_bf, enc_len, dec_len, B, T, Tp):
        m, P, W = self.model, self.P, self.model.weights
        dec, d_enc = m._stacks['decoder'], m._stacks['encoder']['d']
        d, H = dec['d'], dec['heads'][i]
        dh = d // H
        pre = f'decoder.b{i}.'
        rate = self.drop_rate
        c = {'x_f': x_f, 'x_bf': x_bf}
        # ---- self-attention (look-ahead + padding mask), residual, LayerNorm
        qkv = self._bf(B, T, 3 * d)
        m._gemm(P[pre + 'sa.qkv'], B, T, [(x_bf, None, d, 0)], [0], [0], out_hi=qkv, ld_out=3 * d)
        attn, c_sa = self._attn_fwd(B, H, dh, T, T, qkv, 3 * d, 0, qkv, 3 * d, d, 2 * d, dec_len, CAUSAL | FULLQ)
        y_f, y_bf, u1 = self._f32(B, T, d), self._bf(B, T, d), self._f32(B, T, d)
        site_o = self._site()
        m._gemm(P[pre + 'sa.wo'], B, T, [(x_bf, None, d, 0), (attn, None, d, 0)], [0, 1], [0, 0], residual=x_f,
                ln=(W[pre + 'sa.ln.gamma'], W[pre + 'sa.ln.beta']), out_f32=y_f, out_hi=y_bf, out_preln=u1, dropout=(rate, site_o))
        # ---- cross-attention onto the encoder output
        qb = self._bf(B, T, d)
        kvb = self._bf(B, Tp, 2 * d)
        m._gemm(P[pre + 'ca.q'], B, T, [(y_bf, None, d, 0)], [0], [0], out_hi=qb, ld_out=d)
        m._gemm(P[pre + 'ca.kv'], B, Tp, [(enc_bf, None, d_enc, 0)], [0], [0], out_hi=kvb, ld_out=2 * d)
        ca, c_ca = self._attn_fwd(B, H, dh, T, Tp, qb, d, 0, kvb, 2 * d, 0, d, enc_len, FULLQ)
        z_f, z_bf, u2 = self._f32(B, T, d), self._bf(B, T, d), self._f32(B, T, d)
        site_c = self._site()
        m._gemm(P[pre + 'ca.wo'], B, T, [(y_bf, None, d, 0), (ca, None, d, 0)], [0, 1], [0, 0], residual=y_f,
                ln=(W[pre + 'ca.ln.gamma'], W[pre + 'ca.ln.beta']), out_f32=z_f, out_hi=z_bf, out_preln=u2, dropout=(rate, site_c))
        # ---- feed-forward
        F = int(dec['ffn'])
        h = self._bf(B, T, F)
        m._gemm(P[pre + 'ffn1'], B, T, [(z_bf, None, d, 0)], [0], [0], relu=True, out_hi=h, ld_out=F)
        o_f, o_bf, u3 = self._f32(B, T, d), self._bf(B, T, d), self._f32(B, T, d)
        site_f = self._site()
        m._gemm(P[pre + 'ffn2'], B, T, [(h, None, F, 0)], [0], [0], residual=z_f, ln=(W[pre + 'ln2.gamma'], W[pre + 'ln2.beta']),
                out_f32=o_f, out_hi=o_bf, out_preln=u3, dropout=(rate, site_f))
        c.update(qkv=qkv, attn=attn, sa=c_sa, y_bf=y_bf, u1=u1, qb=qb, kvb=kvb, ca=ca, cac=c_ca, z_bf=z_bf, u2=u2, h=h, u3=u3,
                 sites=(site_o, site_c, site_f))
        return o_f, o_bf, c

    def _cadb_bwd(self, i, c, do, enc_bf, enc_len, dec_len, d_enc_acc, B, T, Tp, diag):
        m, P, W, G = self.model, self.P, self.model.weights, self.g
        dec, d_enc = m._stacks['decoder'], m._stacks['encoder']['d']
        d, H = dec['d'], dec['heads'][i]
        dh = d // H
        pre = f'decoder.b{i}.'
        rate = self.drop_rate
        site_o, site_c, site_f = c['sites']
        F = int(dec['ffn'])
        # ---- FFNResNorm
        du3, g3 = self._f32(B, T, d), self._bf(B, T, d)
        lib.layernorm_bwd(do, c['u3'], W[pre + 'ln2.gamma'], B, T, d, d, LN_EPS, None, False, du3, g3, G[pre + 'ln2.gamma'],
                          G[pre + 'ln2.beta'], pre_drop=(rate, site_f), seed=self.seed, dbias=G[pre + 'ffn2.b'])
        h = c['h']
        self._wgrad([(h, F)], g3, d, B, T, F, d, [(0, 0)], G[pre + 'ffn2.w'])
        dh_ = self._bf(B, T, F)
        m._gemm(P[pre + 'ffn2.d'], B, T, [(g3, None, d, 0)], [0], [0], out_hi=dh_, ld_out=F)
        lib.relu_bwd(dh_, h)
        lib.colsum_bf16(dh_, B * T, F, F, G[pre + 'ffn1.b'])
        self._wgrad([(c['z_bf'], d)], dh_, F, B, T, d, F, [(0, 0)], G[pre + 'ffn1.w'])
        dz = self._f32(B, T, d)
        m._gemm(P[pre + 'ffn1.d'], B, T, [(dh_, None, F, 0)], [0], [0], residual=du3, out_f32=dz, ld_out=d)
        # ---- CrossAttentionResnorm
        du2, g2 = self._f32(B, T, d), self._bf(B, T, d)
        lib.layernorm_bwd(dz, c['u2'], W[pre + 'ca.ln.gamma'], B, T, d, d, LN_EPS, None, False, du2, g2, G[pre + 'ca.ln.gamma'],
                          G[pre + 'ca.ln.beta'], pre_drop=(rate, site_c), seed=self.seed, dbias=G[pre + 'ca.wo.b'])
        self._wgrad([(c['y_bf'], d), (c['ca'], d)], g2, d, B, T, d, d, [(0, 0), (1, 0)], G[pre + 'ca.wo.w'])
        dca = self._bf(B, T, d)
        m._gemm(P[pre + 'ca.wo.da'], B, T, [(g2, None, d, 0)], [0], [0], out_hi=dca, ld_out=d)
        dy_acc = self._f32(B, T, d)
        m._gemm(P[pre + 'ca.wo.dx'], B, T, [(g2, None, d, 0)], [0], [0], residual=du2, out_f32=dy_acc, ld_out=d)
        dq = self._bf(B, T, d)
        dkv = self._bf(B, Tp, 2 * d)
        self._attn_bwd(c['cac'], B, H, dh, T, Tp, dca, c['qb'], d, 0, c['kvb'], 2 * d, 0, d, enc_len, dq, d, 0, dkv, 2 * d, 0, d,
                       diag=diag)
        lib.colsum_bf16(dq, B * T, d, d, G[pre + 'ca.wq.b'])
        self._wgrad([(c['y_bf'], d)], dq, d, B, T, d, d, [(0, 0)], G[pre + 'ca.wq.w'])
        dy = self._f32(B, T, d)
        m._gemm(P[pre + 'ca.q.d'], B, T, [(dq, None, d, 0)], [0], [0], residual=dy_acc, out_f32=dy, ld_out=d)
        for n_, nm in enumerate(('wk', 'wv')):
            gs = dkv[..., n_ * d:]
            lib.colsum_bf16(gs, B * Tp, d, 2 * d, G[pre + 'ca.' + nm + '.b'])
            self._wgrad([(enc_bf, d_enc)], gs, 2 * d, B, Tp, d_enc, d, [(0, 0)], G[pre + 'ca.' + nm + '.w'])
        d_enc_new = self._f32(B, Tp, d_enc)
        m._gemm(P[pre + 'ca.kv.d'], B, Tp, [(dkv, None, 2 * d, 0)], [0], [0], residual=d_enc_acc, out_f32=d_enc_new, ld_out=d_enc)
        # ---- SelfAttentionResNorm
        du1, g1 = self._f32(B, T, d), self._bf(B, T, d)
        lib.layernorm_bwd(dy, c['u1'], W[pre + 'sa.ln.gamma'], B, T, d, d, LN_EPS, None, False, du1, g1, G[pre + 'sa.ln.gamma'],
                          G[pre + 'sa.ln.beta'], pre_drop=(rate, site_o), seed=self.seed, dbias=G[pre + 'sa.wo.b'])
        self._wgrad([(c['x_bf'], d), (c['attn'], d)], g1, d, B, T, d, d, [(0, 0), (1, 0)], G[pre + 'sa.wo.w'])
        dattn = self._bf(B, T, d)
        m._gemm(P[pre + 'sa.wo.da'], B, T, [(g1, None, d, 0)], [0], [0], out_hi=dattn, ld_out=d)
        dx_acc = self._f32(B, T, d)
        m._gemm(P[pre + 'sa.wo.dx'], B, T, [(g1, None, d, 0)], [0], [0], residual=du1, out_f32=dx_acc, ld_out=d)
        qkv = c['qkv']
        dqkv = self._bf(B, T, 3 * d)
        self._attn_bwd(c['sa'], B, H, dh, T, T, dattn, qkv, 3 * d, 0, qkv, 3 * d, d, 2 * d, dec_len, dqkv, 3 * d, 0, dqkv, 3 * d, d, 2 * d)
        for n_, nm in enumerate(('wq', 'wk', 'wv')):
            gs = dqkv[..., n_ * d:]
            lib.colsum_bf16(gs, B * T, d, 3 * d, G[pre + 'sa.' + nm + '.b'])
            self._wgrad([(c['x_bf'], d)], gs, 3 * d, B, T, d, d, [(0, 0)], G[pre + 'sa.' + nm + '.w'])
        dx = self._f32(B, T, d)
        m._gemm(P[pre + 'sa.qkv.d'], B, T, [(dqkv, None, 3 * d, 0)], [0], [0], residual=dx_acc, out_f32=dx, ld_out=d)
        return dx, d_enc_new

    # ------------------------------------------------------------------------------------------------
    # full step
    # ------------------------------------------------------------------------------------------------
    def step_graphed(self, inp, tar, stop_prob):
        """forward + backward of the teacher-forced step as ONE CUDA graph per input shape (single process; with a gradient
        all-reduce in the middle the eager path is used).  Per-step dropout masks come from the device-resident salt
        (TrainEngine._set_salt); Adam stays an eager launch."""
        m = self.model
        inp, tar, stop_prob = torch.as_tensor(inp), torch.as_tensor(tar), torch.as_tensor(stop_prob)
        key = (tuple(inp.shape), tuple(tar.shape), int(m.r), m.force_encoder_diagonal, m.force_decoder_diagonal, bool(m.train_dropout))
        ent = self._graphs.get(key)
        if ent is None:
            dev = self.dev
            ins = [inp.to(device=dev, dtype=torch.int32).contiguous().clone(), tar.to(device=dev, dtype=torch.float32).contiguous().clone(),
                   stop_prob.to(device=dev, dtype=torch.int32).contiguous().clone()]
            self._set_salt(1)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.forward_backward(*ins, training=True)
            torch.cuda.current_stream().wait_stream(side)
            if self._graph_pool is None:
                self._graph_pool = torch.cuda.graph_pool_handle()
            g = torch.cuda.CUDAGraph()
            n0 = lib.launch_count()
            with torch.cuda.graph(g, pool=self._graph_pool):
                lib.set_dropout_salt(self._salt_dev)
                out = self.forward_backward(*ins, training=True)
            if len(self._graphs) >= 4:
                self._graphs.pop(next(iter(self._graphs)))
            ent = self._graphs[key] = {'ins': ins, 'g': g, 'out': out, 'n': lib.launch_count() - n0}
        else:
            for dst, src in zip(ent['ins'], (inp, tar, stop_prob)):
                dst.copy_(src, non_blocking=True)
        it = m.optimizer.iterations if m.optimizer else 0
        self._set_salt(((it + 1) * 40503 + 12345) & 0x7fffffff)
        self._salt_applied = 1
        ent['g'].replay()
        lib.add_launch_count(ent['n'])
        out = dict(ent['out'])
        out['loss'] = out['loss'].clone()
        out['losses'] = {k: v.clone() for k, v in out['losses'].items()}
        return out

    def forward_backward(self, inp, tar, stop_prob, training=True, sync=None):
        m, W, G = self.model, self.model.weights, self.g
        dev = self.dev
        self.use_dropout = training and m.train_dropout
        self.drop_rate = float(m.config.get('dropout_rate', 0.0)) if self.use_dropout else 0.0
        prenet_rate = float(m.config.get('decoder_prenet_dropout', 0.0)) if self.use_dropout else 0.0
        self.drop_sites = 0
        self.seed = (self.base_seed * 2654435761 + (m.optimizer.iterations if m.optimizer else 0) * 40503 + self.rank * 97) & 0x7fffffff
        m._drop_seed = self.seed
        saved_precision = m.precision
        m.precision = 'bf16'
        try:
            P = self._pack()
            r = int(m.r)
            mel = m.mel_channels
            kmel = _round_up(mel, 64)
            x = torch.as_tensor(inp).to(device=dev, dtype=torch.int32).contiguous()
            tar = torch.as_tensor(tar).to(device=dev, dtype=torch.float32)
            stop = torch.as_tensor(stop_prob).to(device=dev, dtype=torch.int32)
            tar_inp, tar_real, tar_stop = tar[:, :-1], tar[:, 1:].contiguous(), stop[:, 1:].contiguous()
            mel_len = tar_inp.shape[1]
            tgt = tar_inp[:, 0::r, :].contiguous()
            B, Tp = x.shape
            T = tgt.shape[1]
            d_enc, d = m._stacks['encoder']['d'], m._stacks['decoder']['d']
            self._scratch1 = torch.zeros(1, dtype=torch.float32, device=dev)
            enc_len = torch.empty((B,), dtype=torch.int32, device=dev)
            lib.phoneme_lengths(x, 0, enc_len)
            dec_len = torch.empty((B,), dtype=torch.int32, device=dev)
            lib.mel_lengths(tgt, 0.0, dec_len)
            # ---- encoder
            e_rows = self._f32(1, B * Tp, d_enc)
            lib.length_regulate_fwd(W['embedding'].view(1, -1, d_enc), x.view(1, -1), e_rows)
            h_f, h_bf = self._f32(B, Tp, d_enc), self._bf(B, Tp, d_enc)
            site_e = self._site()
            lib.embed_ln_pe_fwd(x, W['embedding'], W['encoder.ln.gamma'], W['encoder.ln.beta'], P['encoder.pe'],
                                W['encoder.pos_scalar'].reshape(1), LN_EPS, h_f, h_bf, None, drop=(self.drop_rate, self.seed, site_e))
            enc_ctx = []
            for i in range(len(m._stacks['encoder']['heads'])):
                h_f, h_bf, c = self._block_fwd('encoder', i, h_f, h_bf, enc_len, B, Tp)
                enc_ctx.append(c)
            enc_bf = h_bf
            # ---- decoder prenet (layers.py:420-443): relu Dense -> dropout -> relu Dense -> dropout
            t_pad = self._bf(B, T, kmel)
            lib.cast_bf16_pad(tgt, B * T, mel, t_pad, kmel)
            pdim = int(m.config['decoder_prenet_dimension'])
            h1 = self._bf(B, T, pdim)
            site_p1 = self._site()
            m._gemm(P['prenet.d1'], B, T, [(t_pad, None, kmel, 0)], [0], [0], relu=True, out_hi=h1, ld_out=pdim, dropout=(prenet_rate, site_p1))
            pre_out, h2 = self._f32(B, T, d), self._bf(B, T, d)
            site_p2 = self._site()
            m._gemm(P['prenet.d2'], B, T, [(h1, None, pdim, 0)], [0], [0], relu=True, out_f32=pre_out, out_hi=h2, ld_out=d,
                    dropout=(prenet_rate, site_p2))
            # ---- CrossAttentionBlocks prologue: LN(inputs) + scalar * PE[:, :T*r:r] -> dropout
            P['decoder.pe'] = m._decoder_pe(r)
            idx = torch.arange(T, dtype=torch.int32, device=dev)[None, :].expand(B, T).contiguous()
            x_f, x_bf = self._f32(B, T, d), self._bf(B, T, d)
            site_d = self._site()
            lib.expand_ln_pe_fwd(pre_out, idx, W['decoder.ln.gamma'], W['decoder.ln.beta'], P['decoder.pe'],
                                 W['decoder.pos_scalar'].reshape(1), LN_EPS, x_f, x_bf, None, drop=(self.drop_rate, self.seed, site_d))
            dec_ctx = []
            n_dec = len(m._stacks['decoder']['heads'])
            for i in range(n_dec):
                x_f, x_bf, c = self._cadb_fwd(i, x_f, x_bf, enc_bf, enc_len, dec_len, B, T, Tp)
                dec_ctx.append(c)
            # ---- FinalProj[:, :, :r*mel] -> (B, T*r, mel) -> Postnet
            fp, fp_d = self._fp[r]
            n_fp = r * mel
            lin = self._f32(B, T, n_fp)
            m._gemm(fp, B, T, [(x_bf, None, d, 0)], [0], [0], out_f32=lin, ld_out=n_fp)
            Tr = T * r
            linear = lin.view(B, Tr, mel)
            l_pad = self._bf(B, Tr, kmel)
            lib.cast_bf16_pad(linear, B * Tr, mel, l_pad, kmel)
            pn = P['postnet']
            post = self._f32(B, Tr, pn.n_pad)
            m._gemm(pn, B, Tr, [(l_pad, None, kmel, 0)], [0], [0], out_f32=post)
            mel_out = post[..., :mel].contiguous()
            stop_out = post[..., mel:mel + 3].contiguous()
            # ---- losses (models.py:179-207; loss weights [1, 1])
            wts = m.loss_weights
            losses = torch.zeros(3, dtype=torch.float32, device=dev)
            dmel = self._f32(B, Tr, mel)
            dstop = self._f32(B, Tr, 3)
            lib.mae_loss(mel_out, B, Tr, mel_len, mel, tar_real, wts[0], losses[0:1], dmel)
            lib.scaled_ce_loss(stop_out, mel_len, 3, tar_stop, m.stop_prob_index, m.stop_scaling, losses[1:2], wts[1], dstop)
            n_maps = (n_dec if m.force_decoder_diagonal else 0) + (len(enc_ctx) if m.force_encoder_diagonal else 0)
            norm = 1.0 + n_maps
            if m.force_decoder_diagonal:
                for i, c in enumerate(dec_ctx):
                    H = m._stacks['decoder']['heads'][i]
                    lib.diag_loss_train(c['cac']['P_drop'], B, H, T, Tp, c['cac']['ldp'], dec_len, enc_len, 1.0 / norm, losses[2:3], 0.0, None)
            if m.force_encoder_diagonal:
                for i, c in enumerate(enc_ctx):
                    H = m._stacks['encoder']['heads'][i]
                    lib.diag_loss_train(c['P_drop'], B, H, Tp, Tp, c['ldp'], enc_len, enc_len, 1.0 / norm, losses[2:3], 0.0, None)
            out = {'mel': mel_out, 'stop_prob': stop_out, 'linear': linear, 'decoder_output': x_f,
                   'mel_mask': mask_from_lengths(dec_len, T), 'text_mask': mask_from_lengths(enc_len, Tp),
                   'mel_lengths': dec_len, 'text_lengths': enc_len, 'decoder_attention': {}, 'encoder_attention': {},
                   'losses': {'mel': losses[0], 'stop_prob': losses[1], 'diag_loss': losses[2]},
                   'loss': wts[0] * losses[0] + wts[1] * losses[1] + losses[2]}
            if not training:
                return out
            # =============================== backward ===============================
            self.flat_g.zero_()
            g_post = self._bf(B, Tr, kmel)   # columns 0..79: d mel, 80..82: d stop logits, rest zero
            gp32 = torch.zeros((B, Tr, kmel), dtype=torch.float32, device=dev)
            gp32[..., :mel] = dmel
            gp32[..., mel:mel + 3] = dstop
            lib.cast_bf16_pad(gp32, B * Tr, kmel, g_post, kmel)
            lib.colsum_bf16(g_post, B * Tr, mel, kmel, G['postnet.mel.b'])
            lib.colsum_bf16(g_post[..., mel:], B * Tr, 3, kmel, G['postnet.stop.b'])
            self._wgrad([(l_pad, kmel)], g_post, kmel, B, Tr, mel, mel, [(0, 0)], G['postnet.mel.w'])
            self._wgrad([(l_pad, kmel)], g_post[..., mel:], kmel, B, Tr, mel, 3, [(0, 0)], G['postnet.stop.w'])
            dlin = self._f32(B, Tr, mel)
            m._gemm(P['postnet.d'], B, Tr, [(g_post, None, kmel, 0)], [0], [0], out_f32=dlin, ld_out=mel)
            # FinalProj: (B, T*r, mel) gradient is the (B, T, r*mel) gradient of the sliced Dense output
            nfp_pad = _round_up(n_fp, 64)
            g_fp = self._bf(B, T, nfp_pad)
            lib.cast_bf16_pad(dlin.view(B * T, n_fp), B * T, n_fp, g_fp, nfp_pad)
            db = torch.zeros(_round_up(n_fp, 8), dtype=torch.float32, device=dev)
            lib.colsum_bf16(g_fp, B * T, n_fp, nfp_pad, db)
            G['final_proj.b'][:n_fp].add_(db[:n_fp])
            dw = torch.zeros((d, n_fp), dtype=torch.float32, device=dev)
            self._wgrad([(x_bf, d)], g_fp, nfp_pad, B, T, d, n_fp, [(0, 0)], dw)
            G['final_proj.w'][:, :n_fp].add_(dw)
            dz = self._f32(B, T, d)
            m._gemm(fp_d, B, T, [(g_fp, None, nfp_pad, 0)], [0], [0], out_f32=dz, ld_out=d)
            # decoder blocks
            d_enc_acc = None
            diag = (1.0 / norm, dec_len, enc_len) if m.force_decoder_diagonal else None
            for i in range(n_dec - 1, -1, -1):
                dz, d_enc_acc = self._cadb_bwd(i, dec_ctx[i], dz, enc_bf, enc_len, dec_len, d_enc_acc, B, T, Tp, diag)
                dec_ctx[i] = None
            d_pre = self._prologue_bwd('decoder', dz, pre_out, None, B, T, site_d)
            # prenet: relu + dropout gradients from the saved (post-dropout) outputs, then the two Dense layers
            keep = 1.0 / (1.0 - prenet_rate) if prenet_rate > 0 else 1.0
            g2 = self._bf(B, T, d)
            lib.cast_bf16_pad(d_pre, B * T, d, g2, d)
            lib.relu_bwd(g2, h2)
            if keep != 1.0:
                g2.mul_(keep)
            lib.colsum_bf16(g2, B * T, d, d, G['prenet.d2.b'])
            self._wgrad([(h1, pdim)], g2, d, B, T, pdim, d, [(0, 0)], G['prenet.d2.w'])
            g1 = self._bf(B, T, pdim)
            m._gemm(P['prenet.d2.d'], B, T, [(g2, None, d, 0)], [0], [0], out_hi=g1, ld_out=pdim)
            lib.relu_bwd(g1, h1)
            if keep != 1.0:
                g1.mul_(keep)
            lib.colsum_bf16(g1, B * T, pdim, pdim, G['prenet.d1.b'])
            self._wgrad([(t_pad, kmel)], g1, pdim, B, T, mel, pdim, [(0, 0)], G['prenet.d1.w'])
            if sync is not None:
                sync.bucket_ready(*self.decoder_range)
            # encoder (the encoder maps' diagonal loss adds to dP inside the block backward)
            dz = d_enc_acc
            self._enc_diag = (1.0 / norm, enc_len) if m.force_encoder_diagonal else None
            for i in range(len(enc_ctx) - 1, -1, -1):
                dz = self._block_bwd('encoder', i, enc_ctx[i], dz, enc_len, B)
                enc_ctx[i] = None
            de = self._prologue_bwd('encoder', dz, e_rows.view(B, Tp, d_enc), enc_len, B, Tp, site_e)
            lib.embedding_bwd(de, x, G['embedding'])
            return out
        finally:
            m.precision = saved_precision
