"""Host side of the C ABI: packs a reference state_dict into the device layout the kernels want, owns the
gam_handle and the scratch workspace, and turns torch tensors into raw pointers.  No arithmetic of the path is
done here -- only one-off weight re-layout at load time (BatchNorm folding, q/k concatenation, GLU row pairing,
conv weight permutation, fp16 casts, DFT / rotary tables)."""
from __future__ import annotations

import collections
import ctypes as C
import hashlib
import json
import math
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib

Tensor = torch.Tensor


def _cfg_get(section, key, default=None):
    if isinstance(section, dict):
        return section.get(key, default)
    return getattr(section, key, default) if hasattr(section, key) else (section.get(key, default) if hasattr(section, "get") else default)


# ---------------------------------------------------------------------------------- pure weight re-layout (CPU-testable)
def glu_row_permutation(d: int, half: int = 128) -> Tensor:
    """Row order of pointwise_conv1 so that accumulator tile j (2*half columns) = [value rows j*half.. | gate rows d+j*half..]:
    the GEMM epilogue then computes value * sigmoid(gate) thread-locally (gigaam/encoder.py:398-399 GLU over channels)."""
    return torch.cat([torch.cat([torch.arange(j * half, (j + 1) * half), d + torch.arange(j * half, (j + 1) * half)])
                      for j in range(d // half)])


def fold_batchnorm(dw_w: Tensor, dw_b: Tensor, gamma: Tensor, beta: Tensor, mean: Tensor, var: Tensor, eps: float = 1e-5):
    """Eval-mode BatchNorm1d after the depthwise conv folded into its weights (gigaam/encoder.py:402-405)."""
    s = gamma / torch.sqrt(var + eps)
    return dw_w * s[:, None], (dw_b - mean) * s + beta


def pack_conv2_weight(w2: Tensor) -> Tensor:
    """[C_out, C_in, kt, kf] -> [C_out, (kt, kf, C_in)]: K order of the implicit GEMM (tap-major, channel-minor)."""
    return w2.permute(0, 2, 3, 1).reshape(w2.shape[0], -1)


def pack_sub_out_weight(wo: Tensor, channels: int) -> Tensor:
    """pre_encode.out.weight [d, C*F2] with K index c*F2+f (gigaam/encoder.py:125-127) -> K index f*C+c, the order in
    which the stage-2 conv epilogue writes its [B, T', F2, C] output."""
    f2 = wo.shape[1] // channels
    return wo.reshape(wo.shape[0], channels, f2).permute(0, 2, 1).reshape(wo.shape[0], f2 * channels)


DFT_BASIS_SCALE = 8.0      # must match kBasisScale / kFrameScale in csrc/gam_api.cu, frontend.cu
DFT_FRAME_SCALE = 2048.0


def split_dft_basis(n_fft: int) -> Tensor:
    """fp16 [512, 3*Kp] basis of the real DFT for the K-concatenated split-precision GEMM (Kp = n_fft rounded up to 64).
    Rows: two 256-row tiles, tile t = [128 cos rows | 128 sin rows] of bins t*128 + j (bins >= n_fft/2+1 are zero rows);
    columns: [d_hi | d_hi | d_lo] with d = cos / sin(2 pi k i / n_fft), d_hi = fp16(d), d_lo = fp16(d - d_hi)."""
    kp = (n_fft + 63) // 64 * 64
    nb = n_fft // 2 + 1
    k = torch.arange(256, dtype=torch.float64)[:, None]
    i = torch.arange(kp, dtype=torch.float64)[None, :]
    ang = 2.0 * math.pi * k * i / n_fft
    valid = ((k < nb) & (i < n_fft)).double()
    # x 8: keeps the fp16 `lo` halves of small basis values out of the subnormal range (the frames are scaled by
    # 2^11 for the same reason; the power epilogue multiplies by 2^-28)
    basis = torch.stack([torch.cos(ang) * valid, torch.sin(ang) * valid], 1) * DFT_BASIS_SCALE   # [256 bins, 2, kp]
    rows = basis.view(2, 128, 2, kp).permute(0, 2, 1, 3).reshape(512, kp)          # tile, (cos|sin), bin-in-tile
    hi = rows.to(torch.float16)
    lo = (rows - hi.double()).to(torch.float16)
    return torch.cat([hi, hi, lo], dim=1).contiguous()


def rel_pos_embedding(max_t: int, d: int) -> Tensor:
    """Sinusoids of the relative positions max_t-1 ... -(max_t-1), row max_t-1-r for position r, sin on even / cos on
    odd columns (gigaam/encoder.py:318-326).  The reference slices the same rows out of its pos_emb_max_len table
    (:329-334), so the projected table below serves every T' <= max_t."""
    pos = torch.arange(max_t - 1, -max_t, -1, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe = torch.zeros(2 * max_t - 1, d)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def pack_rel_pos_qkv(wq: Tensor, bq: Tensor, wk: Tensor, bk: Tensor, wv: Tensor, bv: Tensor, bias_u: Tensor, bias_v: Tensor):
    """One projection for the rel_pos attention: rows [q ; q ; k ; v] with pos_bias_u / pos_bias_v (gigaam/encoder.py:
    221-222, [h, d_k] = the d_model axis split by head) folded into the two q biases -> ([4d, d], [4d])."""
    w = torch.cat([wq, wq, wk, wv], 0)
    b = torch.cat([bq + bias_u.reshape(-1), bq + bias_v.reshape(-1), bk, bv], 0)
    return w, b


def rotary_half_tables(dk: int, base: float, max_len: int):
    """cos/sin [max_len, dk/2] of t * base^(-2i/dk) (gigaam/encoder.py:342-355; base = pos_emb_max_len)."""
    inv_freq = 1.0 / (base ** (torch.arange(0, dk, 2).float() / dk))
    freqs = torch.einsum("i,j->ij", torch.arange(max_len).float(), inv_freq)
    return freqs.cos(), freqs.sin()


class _WorkspaceCache:
    """Bounded LRU of scratch tensors of ONE kind (encode / log-mel / decode).  Eviction only drops this cache's
    reference: a CUDA graph that baked a workspace pointer keeps the tensor alive through `Engine.held_workspaces`."""

    def __init__(self, cap: int):
        self.cap = cap
        self._d: "collections.OrderedDict[Tuple, Tensor]" = collections.OrderedDict()

    def get(self, key, nbytes: int, device) -> Tensor:
        ws = self._d.get(key)
        if ws is not None and ws.numel() >= nbytes:
            self._d.move_to_end(key)
            return ws
        while len(self._d) >= self.cap:
            self._d.popitem(last=False)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self._d[key] = ws
        return ws

    def peek(self, key) -> Optional[Tensor]:
        return self._d.get(key)

    def tensors(self) -> List[Tensor]:
        return list(self._d.values())

    def __len__(self):
        return len(self._d)


class Engine:
    """One model replica on one CUDA device."""

    WS_CACHE = 4      # workspaces kept per kind (distinct batch shapes)

    PACK_FORMAT = 3   # bump when the packing order / layouts below change

    def __init__(self, cfg: Dict, state_dict: Dict[str, Tensor], device: torch.device, pack_cache: Optional[str] = None):
        """`pack_cache`: path of an on-disk cache of the re-laid-out weights.  When it exists and matches
        this cfg the load-time re-layout (BN folding, concatenations, permutations, fp16 casts, tables) is skipped and the
        packed tensors are uploaded as stored; otherwise it is written after packing.  Callers key the path by the
        checkpoint's md5 (load_model does)."""
        if device.type != "cuda":
            raise RuntimeError("gigaam_b200 runs on CUDA (sm_100a) devices only; there is no CPU path")
        if device.index is None:      # an index-less "cuda" means the CURRENT device, not GPU 0
            device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.load()
        self.device = device
        self.cfg = cfg
        self._keep: List[Tensor] = []
        self._pack_sig = hashlib.sha256(json.dumps([self.PACK_FORMAT, cfg], sort_keys=True, default=str).encode()).hexdigest()
        self._pack_in: Optional[List[Tensor]] = None      # tensors replayed from the cache, in _dev() call order
        self._pack_out: Optional[List[Tensor]] = None     # tensors recorded for the cache
        if pack_cache is not None:
            self._pack_in = self._read_pack_cache(pack_cache)
            if self._pack_in is None:
                self._pack_out = []
        self._ws_enc = _WorkspaceCache(self.WS_CACHE)
        self._ws_mel = _WorkspaceCache(self.WS_CACHE)
        self._ws_dec = _WorkspaceCache(self.WS_CACHE)
        self.handle = C.c_void_p()
        pre, enc = cfg["preprocessor"], cfg["encoder"]
        head = cfg.get("head") if isinstance(cfg, dict) else None
        sr = _cfg_get(pre, "sample_rate")
        self.n_fft = _cfg_get(pre, "n_fft", sr // 40)
        self.win = _cfg_get(pre, "win_length", sr // 40)
        self.hop = _cfg_get(pre, "hop_length", sr // 100)
        self.center = bool(_cfg_get(pre, "center", True))
        self.n_mels = _cfg_get(pre, "features")
        self.d_model = enc["d_model"]
        self.n_layers = enc["n_layers"]
        self.n_heads = enc["n_heads"]
        self.d_ff = self.d_model * enc["ff_expansion_factor"]
        if enc["self_attention_model"] not in ("rotary", "rel_pos"):
            raise ValueError(f"unknown self_attention_model {enc['self_attention_model']!r}")
        self.rel_pos = enc["self_attention_model"] == "rel_pos"
        self._pos_emb = rel_pos_embedding(_lib.REL_POS_MAX_T, self.d_model).to(device) if self.rel_pos else None
        self.head_type = 0
        self.num_classes = 0
        self.max_symbols = 10
        gc = _lib.GamConfig()
        gc.sample_rate, gc.n_mels, gc.n_fft, gc.win_length, gc.hop_length, gc.center = sr, self.n_mels, self.n_fft, self.win, self.hop, int(self.center)
        gc.feat_in, gc.n_layers, gc.d_model, gc.n_heads, gc.d_ff = enc["feat_in"], self.n_layers, self.d_model, self.n_heads, self.d_ff
        gc.subsampling = 0 if enc["subsampling"] == "conv2d" else 1
        gc.subs_kernel_size = enc["subs_kernel_size"]
        gc.conv_kernel_size = enc["conv_kernel_size"]
        gc.conv_norm = 0 if enc["conv_norm_type"] == "batch_norm" else 1
        gc.self_attention = 1 if self.rel_pos else 0
        gc.pos_emb_max_len = enc["pos_emb_max_len"]
        gw = _lib.GamWeights()
        sd = state_dict
        self._pack_frontend(gw, sd)
        self._pack_subsampling(gw, sd, enc)
        self._pack_rope(gw, enc)
        self._layers = (_lib.GamLayerWeights * self.n_layers)()
        for l in range(self.n_layers):
            self._pack_layer(self._layers[l], sd, l, enc)
        gw.layers = C.cast(self._layers, C.POINTER(_lib.GamLayerWeights))
        if head is not None:
            if head["type"] == "ctc":
                self.head_type, self.num_classes = 1, head["num_classes"]
                gw.ctc_w = self._dev(sd["head.decoder_layers.0.weight"].reshape(self.num_classes, -1).float())
                gw.ctc_b = self._dev(sd["head.decoder_layers.0.bias"].float())
            else:
                self._pack_rnnt(gw, gc, sd, head)
                self.max_symbols = int(_cfg_get(cfg.get("decoding", {}), "max_symbols_per_step", 10))
        gc.head, gc.num_classes, gc.max_symbols = self.head_type, self.num_classes, self.max_symbols
        with torch.cuda.device(device):
            rc = self.lib.gam_create(C.byref(gc), C.byref(gw), device.index, C.byref(self.handle))
        _lib.check(self.lib, self.handle, rc, "gam_create")
        if self._pack_out is not None:
            self._write_pack_cache(pack_cache)
        self.pack_cache_hit = self._pack_in is not None
        self._pack_in = self._pack_out = None

    # ------------------------------------------------------------------ packing helpers
    def _dev(self, t, dtype: Optional[torch.dtype] = None) -> int:
        """Upload one packed tensor (or, given a callable, the tensor it computes) and keep it alive; returns the device
        pointer.  With a cache hit the stored tensor of this call position is uploaded and `t` is never evaluated."""
        if self._pack_in is not None:
            t = self._pack_in[len(self._keep)]
        else:
            t = (t() if callable(t) else t).detach()
            if dtype is not None:
                t = t.to(dtype)
            if self._pack_out is not None:
                self._pack_out.append(t.cpu().contiguous())
        t = t.to(self.device).contiguous()
        self._keep.append(t)
        return t.data_ptr()

    def _read_pack_cache(self, path: str) -> Optional[List[Tensor]]:
        if not os.path.isfile(path):
            return None
        try:
            blob = torch.load(path, map_location="cpu", weights_only=True)
            if blob.get("signature") == self._pack_sig and isinstance(blob.get("tensors"), list):
                return blob["tensors"]
        except Exception:
            pass
        return None          # stale or unreadable: repack and overwrite

    def _write_pack_cache(self, path: str) -> None:
        try:
            tmp = f"{path}.tmp{os.getpid()}"
            torch.save({"signature": self._pack_sig, "tensors": self._pack_out}, tmp)
            os.replace(tmp, path)
        except OSError:
            pass             # a read-only cache directory must not break loading

    def _pack_frontend(self, gw, sd):
        n = self.n_fft
        K = n // 2 + 1
        window = sd["preprocessor.featurizer.0.spectrogram.window"].float()
        if window.numel() != n:
            raise NotImplementedError("win_length != n_fft")
        idx = torch.arange(K, dtype=torch.float64)
        ang = 2.0 * math.pi * torch.outer(idx, idx) / n  # [n, k]
        gw.window = self._dev(window)
        gw.dft_cos = self._dev(torch.cos(ang).float())
        gw.dft_sin = self._dev(torch.sin(ang).float())
        fb = sd["preprocessor.featurizer.0.mel_scale.fb"].float().cpu()
        gw.mel_fb = self._dev(fb)
        # tensor-core front end: split-precision DFT basis + bin range of every mel filter
        if K <= 256:
            gw.dft_w = self._dev(lambda: split_dft_basis(n))
            nz = fb != 0
            lo = torch.where(nz.any(0), nz.float().argmax(0), torch.zeros(fb.shape[1], dtype=torch.long))
            hi = torch.where(nz.any(0), fb.shape[0] - nz.flip(0).float().argmax(0), torch.zeros(fb.shape[1], dtype=torch.long))
            gw.mel_lo = self._dev(lo.to(torch.int32))
            gw.mel_hi = self._dev(hi.to(torch.int32))
            self._logmel_tc = True
        else:
            self._logmel_tc = False

    def _pack_subsampling(self, gw, sd, enc):
        p = "encoder.pre_encode."
        d = self.d_model
        if enc["subsampling"] == "conv1d":
            # Conv1d weights [out, in, k] -> (out, k, in): K order (tap, channel) of the implicit GEMM
            for i, name in ((0, "c1d_w1"), (2, "c1d_w2")):
                w = sd[f"{p}conv.{i}.weight"].float()
                setattr(gw, name, self._dev(w.permute(0, 2, 1).reshape(w.shape[0], -1), torch.float16))
            gw.c1d_b1 = self._dev(sd[p + "conv.0.bias"].float())
            gw.c1d_b2 = self._dev(sd[p + "conv.2.bias"].float())
            return
        w1 = sd[p + "conv.0.weight"].float()                     # [C, 1, 3, 3]
        gw.sub1_w = self._dev(w1.reshape(d, 9))
        gw.sub1_b = self._dev(sd[p + "conv.0.bias"].float())
        w2 = sd[p + "conv.2.weight"].float()                     # [C_out, C_in, kt, kf]
        gw.sub2_w = self._dev(lambda: pack_conv2_weight(w2), torch.float16)
        gw.sub2_b = self._dev(sd[p + "conv.2.bias"].float())
        wo = sd[p + "out.weight"].float()                        # [d, C*F2] with K index c*F2 + f
        gw.sub_out_w = self._dev(lambda: pack_sub_out_weight(wo, d), torch.float16)
        gw.sub_out_b = self._dev(sd[p + "out.bias"].float())

    def _pack_rope(self, gw, enc):
        dk = self.d_model // self.n_heads
        base = enc["pos_emb_max_len"]  # the reference passes pos_emb_max_len as the rotary base (encoder.py:546-548)
        cos, sin = rotary_half_tables(dk, base, enc["pos_emb_max_len"])
        gw.rope_cos = self._dev(cos)
        gw.rope_sin = self._dev(sin)

    def _pack_layer(self, lw, sd, l: int, enc):
        q = f"encoder.layers.{l}."
        d = self.d_model
        h16 = torch.float16

        def f(name):
            return sd[q + name].float()

        lw.ln_ff1_g, lw.ln_ff1_b = self._dev(f("norm_feed_forward1.weight")), self._dev(f("norm_feed_forward1.bias"))
        lw.ff1_w1, lw.ff1_b1 = self._dev(f("feed_forward1.linear1.weight"), h16), self._dev(f("feed_forward1.linear1.bias"))
        lw.ff1_w2, lw.ff1_b2 = self._dev(f("feed_forward1.linear2.weight"), h16), self._dev(f("feed_forward1.linear2.bias"))
        lw.ln_att_g, lw.ln_att_b = self._dev(f("norm_self_att.weight")), self._dev(f("norm_self_att.bias"))
        if self.rel_pos:
            w4, b4 = pack_rel_pos_qkv(f("self_attn.linear_q.weight"), f("self_attn.linear_q.bias"),
                                      f("self_attn.linear_k.weight"), f("self_attn.linear_k.bias"),
                                      f("self_attn.linear_v.weight"), f("self_attn.linear_v.bias"),
                                      f("self_attn.pos_bias_u"), f("self_attn.pos_bias_v"))
            lw.w_qkv_rel, lw.b_qkv_rel = self._dev(w4, h16), self._dev(b4)
            # linear_pos (no bias, encoder.py:198,219) of a constant table is a constant: projected once at load time
            lw.pos_proj = self._dev(lambda: self._pos_emb @ f("self_attn.linear_pos.weight").to(self.device).t(), h16)
            lw.w_qk = lw.b_qk = lw.w_v = lw.b_v = None
        else:
            # [W_q ; W_k ; W_v] and their biases in ONE allocation each: w_v / b_v point behind w_qk / b_qk, which lets the
            # library run the three projections as a single launch (two A operands: rope(u) for q, k and u for v)
            w_qkv = self._dev(torch.cat([f("self_attn.linear_q.weight"), f("self_attn.linear_k.weight"),
                                         f("self_attn.linear_v.weight")], 0), h16)
            b_qkv = self._dev(torch.cat([f("self_attn.linear_q.bias"), f("self_attn.linear_k.bias"), f("self_attn.linear_v.bias")], 0))
            lw.w_qk, lw.w_v = w_qkv, w_qkv + 2 * d * d * 2
            lw.b_qk, lw.b_v = b_qkv, b_qkv + 2 * d * 4
            lw.w_qkv_rel = lw.b_qkv_rel = lw.pos_proj = None
        lw.w_o, lw.b_o = self._dev(f("self_attn.linear_out.weight"), h16), self._dev(f("self_attn.linear_out.bias"))
        lw.ln_conv_g, lw.ln_conv_b = self._dev(f("norm_conv.weight")), self._dev(f("norm_conv.bias"))
        # GLU pairing: accumulator tile j (256 columns) = [value rows j*128.. | gate rows d + j*128..]
        perm = glu_row_permutation(d)
        w1 = f("conv.pointwise_conv1.weight").reshape(2 * d, d)
        lw.pw1_w, lw.pw1_b = self._dev(w1[perm], h16), self._dev(f("conv.pointwise_conv1.bias")[perm])
        dw = f("conv.depthwise_conv.weight").reshape(d, -1)
        db = f("conv.depthwise_conv.bias")
        if enc["conv_norm_type"] == "batch_norm":
            dw, db = fold_batchnorm(dw, db, f("conv.batch_norm.weight"), f("conv.batch_norm.bias"),
                                    f("conv.batch_norm.running_mean"), f("conv.batch_norm.running_var"))
            lw.cn_g, lw.cn_b = None, None
        else:
            lw.cn_g, lw.cn_b = self._dev(f("conv.batch_norm.weight")), self._dev(f("conv.batch_norm.bias"))
        lw.dw_w, lw.dw_b = self._dev(dw.t()), self._dev(db)     # taps transposed to [k, d]: coalesced per-tap loads
        lw.pw2_w = self._dev(f("conv.pointwise_conv2.weight").reshape(d, d), h16)
        lw.pw2_b = self._dev(f("conv.pointwise_conv2.bias"))
        lw.ln_ff2_g, lw.ln_ff2_b = self._dev(f("norm_feed_forward2.weight")), self._dev(f("norm_feed_forward2.bias"))
        lw.ff2_w1, lw.ff2_b1 = self._dev(f("feed_forward2.linear1.weight"), h16), self._dev(f("feed_forward2.linear1.bias"))
        lw.ff2_w2, lw.ff2_b2 = self._dev(f("feed_forward2.linear2.weight"), h16), self._dev(f("feed_forward2.linear2.bias"))
        lw.ln_out_g, lw.ln_out_b = self._dev(f("norm_out.weight")), self._dev(f("norm_out.bias"))

    def _pack_rnnt(self, gw, gc, sd, head):
        dc, jt = head["decoder"], head["joint"]
        if dc["pred_rnn_layers"] != 1:
            raise NotImplementedError("multi-layer prediction LSTM")
        self.head_type, self.num_classes = 2, jt["num_classes"]
        gc.pred_hidden, gc.joint_hidden = dc["pred_hidden"], jt["joint_hidden"]
        emb = sd["head.decoder.embed.weight"].double().clone()
        # predict(None, None) starts from an all-zero embedding (gigaam/decoder.py:92-95) and nn.Embedding's padding_idx
        # row is zero by construction but not enforced by load_state_dict: the table's blank row carries the biases only
        emb[jt["num_classes"] - 1].zero_()
        w_ih, w_hh = sd["head.decoder.lstm.weight_ih_l0"].double(), sd["head.decoder.lstm.weight_hh_l0"].float()
        bias = sd["head.decoder.lstm.bias_ih_l0"].double() + sd["head.decoder.lstm.bias_hh_l0"].double()
        gw.rnnt_emb_gates = self._dev((emb @ w_ih.t() + bias).float())
        gw.rnnt_whh_t = self._dev(w_hh.t())
        gw.rnnt_wp_t = self._dev(sd["head.joint.pred.weight"].float().t())
        gw.rnnt_bp = self._dev(sd["head.joint.pred.bias"].float())
        gw.rnnt_enc_w = self._dev(sd["head.joint.enc.weight"].float())
        gw.rnnt_enc_b = self._dev(sd["head.joint.enc.bias"].float())
        gw.rnnt_wo = self._dev(sd["head.joint.joint_net.1.weight"].float())
        gw.rnnt_bo = self._dev(sd["head.joint.joint_net.1.bias"].float())

    # ------------------------------------------------------------------ calls
    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def logmel_frames(self, n: int) -> int:
        return int(self.lib.gam_logmel_frames(self.handle, int(n)))

    def encoded_frames(self, m: int) -> int:
        return int(self.lib.gam_encoded_frames(self.handle, int(m)))

    def workspace(self, B: int, M: int) -> Tensor:
        return self._ws_enc.get((B, M), int(self.lib.gam_workspace_bytes(self.handle, B, M)), self.device)

    def held_workspaces(self, B: int, N: int) -> List[Tensor]:
        """The scratch tensors a step over a [B, N] waveform batch touches.  A captured CUDA graph stores this list: the
        pointers it baked stay valid however many other shapes pass through the engine afterwards."""
        M = self.logmel_frames(N)
        T = self.encoded_frames(M)
        held = [self._ws_mel.peek((B, N)), self._ws_enc.peek((B, M)), self._ws_dec.peek((B, T))]
        return [t for t in held if t is not None]

    def logmel(self, wav: Tensor, fused: bool = False) -> Tensor:
        """[B, N] f32 on device -> [B, n_mels, M] f32 (FeatureExtractor.forward, gigaam/preprocess.py:94-98).
        `fused=True` selects the single CUDA-core kernel (no workspace) instead of the tensor-core DFT."""
        assert wav.is_cuda and wav.dtype == torch.float32 and wav.dim() == 2
        wav = wav.contiguous()
        B, N = wav.shape
        M = self.logmel_frames(N)
        mel = torch.empty((B, self.n_mels, M), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            if self._logmel_tc and not fused:
                ws = self._ws_mel.get((B, N), int(self.lib.gam_logmel_workspace_bytes(self.handle, B, N)), self.device)
                rc = self.lib.gam_logmel_tc(self.handle, wav.data_ptr(), B, N, mel.data_ptr(), ws.data_ptr(), ws.numel(), self._stream())
                _lib.check(self.lib, self.handle, rc, "gam_logmel_tc")
                return mel
            rc = self.lib.gam_logmel(self.handle, wav.data_ptr(), B, N, mel.data_ptr(), self._stream())
        _lib.check(self.lib, self.handle, rc, "gam_logmel")
        return mel

    def encode(self, mel: Tensor, mel_len: Tensor, n_layers_run: int = -1) -> Tuple[Tensor, Tensor]:
        """[B, F, M] f32, [B] i64 -> ([B, T', d] f32 row-major, [B] i32)"""
        assert mel.is_cuda and mel.dtype == torch.float32 and mel.dim() == 3
        mel = mel.contiguous()
        mel_len = mel_len.to(device=self.device, dtype=torch.int64).contiguous()
        B, _, M = mel.shape
        T = self.encoded_frames(M)
        ws = self.workspace(B, M)
        enc = torch.empty((B, T, self.d_model), dtype=torch.float32, device=self.device)
        enc_len = torch.empty((B,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.gam_encode(self.handle, mel.data_ptr(), mel_len.data_ptr(), B, M, ws.data_ptr(), ws.numel(),
                                     enc.data_ptr(), enc_len.data_ptr(), n_layers_run, self._stream())
        _lib.check(self.lib, self.handle, rc, "gam_encode")
        return enc, enc_len

    def hyp_width(self, T: int) -> int:
        """Row pitch of the id / frame matrices for T encoder frames."""
        return T if self.head_type == 1 else T * self.max_symbols

    def packed_hypotheses(self, rows: int, T: int) -> Tensor:
        """Zeroed int32 buffer [ids rows x W | frames rows x W | counts rows] (the layout gam_gather_hyps all-gathers)."""
        w = self.hyp_width(T)
        return torch.zeros(2 * rows * w + rows, dtype=torch.int32, device=self.device)

    def greedy(self, enc_btd: Tensor, enc_len: Tensor, packed: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
        """enc [B, T, d] f32 contiguous, len [B] -> (ids [B, max_out] i32, frames, counts [B] i32) on device.
        `packed` (from packed_hypotheses, rows >= B): the results are written into that buffer and returned as views of it."""
        assert enc_btd.is_cuda and enc_btd.dtype == torch.float32 and enc_btd.is_contiguous()
        if self.head_type == 0:
            raise RuntimeError("model has no head to decode with")
        B, T, _ = enc_btd.shape
        enc_len = enc_len.to(device=self.device, dtype=torch.int32).contiguous()
        max_out = self.hyp_width(T)
        if packed is not None:
            rows = packed.numel() // (2 * max_out + 1)
            assert rows >= B and packed.numel() == rows * (2 * max_out + 1) and packed.dtype == torch.int32
            ids = packed[: rows * max_out].view(rows, max_out)[:B]
            frames = packed[rows * max_out: 2 * rows * max_out].view(rows, max_out)[:B]
            counts = packed[2 * rows * max_out: 2 * rows * max_out + B]
        else:
            ids = torch.empty((B, max_out), dtype=torch.int32, device=self.device)
            frames = torch.empty((B, max_out), dtype=torch.int32, device=self.device)
            counts = torch.empty((B,), dtype=torch.int32, device=self.device)
        ws = self._ws_dec.get((B, T), int(self.lib.gam_decode_workspace_bytes(self.handle, B, T)), self.device)
        fn = self.lib.gam_ctc_greedy if self.head_type == 1 else self.lib.gam_rnnt_greedy
        with torch.cuda.device(self.device):
            rc = fn(self.handle, enc_btd.data_ptr(), enc_len.data_ptr(), B, T, ws.data_ptr(), ws.numel(), ids.data_ptr(),
                    frames.data_ptr(), counts.data_ptr(), max_out, self._stream())
        _lib.check(self.lib, self.handle, rc, "gam_greedy")
        return ids, frames, counts

    def group_words(self, ids: Tensor, frames: Tensor, counts: Tensor, token_flags: Tensor):
        """Device word grouping (gam_group_words): -> (word_start, word_end, word_first, word_ntok [B, max_out] i32, n_words [B] i32)."""
        B, max_out = ids.shape
        flags = token_flags.to(device=self.device, dtype=torch.uint8).contiguous()
        outs = [torch.empty((B, max_out), dtype=torch.int32, device=self.device) for _ in range(4)]
        n_words = torch.empty((B,), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.gam_group_words(self.handle, ids.data_ptr(), frames.data_ptr(), counts.data_ptr(), B, max_out, flags.data_ptr(),
                                          flags.numel(), max_out, *[t.data_ptr() for t in outs], n_words.data_ptr(), self._stream())
        _lib.check(self.lib, self.handle, rc, "gam_group_words")
        return (*outs, n_words)

    def profile_begin(self) -> None:
        _lib.check(self.lib, self.handle, self.lib.gam_profile_begin(self.handle), "gam_profile_begin")

    def profile_end(self) -> Dict[str, Tuple[float, int]]:
        """{kernel class: (total ms, launches)} measured with CUDA events since profile_begin()."""
        n = int(self.lib.gam_profile_class_count())
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        _lib.check(self.lib, self.handle, self.lib.gam_profile_end(self.handle, ms, cnt, n), "gam_profile_end")
        return {self.lib.gam_profile_class_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i] > 0}

    def launch_count(self) -> int:
        return int(self.lib.gam_launch_count(self.handle))

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self.lib.gam_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass
