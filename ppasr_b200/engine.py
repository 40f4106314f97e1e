"""Thin Python handle over the C-ABI context (include/ppasr_b200.h).

torch is used only as a container for device memory and for streams; all compute happens in
libppasr_b200.so. There is no CPU path: constructing an engine without a B200 raises.
"""
import ctypes

import numpy as np

from . import _lib as L
from .weights import ConformerConfig


class Config(ctypes.Structure):
    _fields_ = [("model_type", ctypes.c_int32), ("feat_dim", ctypes.c_int32), ("d_model", ctypes.c_int32),
                ("n_heads", ctypes.c_int32), ("ffn_dim", ctypes.c_int32), ("n_layers", ctypes.c_int32),
                ("conv_kernel", ctypes.c_int32), ("causal", ctypes.c_int32), ("conv_norm", ctypes.c_int32),
                ("vocab_size", ctypes.c_int32), ("max_len", ctypes.c_int32), ("reduce_idx", ctypes.c_int32),
                ("recover_idx", ctypes.c_int32), ("time_reduce_kernel", ctypes.c_int32), ("use_gru", ctypes.c_int32), ("stride_layer_idx", ctypes.c_int32),
                ("group_layer_mask", ctypes.c_int32), ("group_size", ctypes.c_int32), ("stride_kernel", ctypes.c_int32),
                ("reserved", ctypes.c_int32 * 1)]


def out_frames(T: int) -> int:
    """Encoder output frames for T fbank frames (two k3/s2 convs, conformer/subsampling.py:96-115)."""
    return ((T - 1) // 2 - 1) // 2 if T >= 7 else 0


class ConformerEngine:
    """Owns one ppasr_b200_ctx: packed weights, workspace and (later) streaming caches."""

    def __init__(self, cfg: ConformerConfig, weights, device: int = 0):
        import torch
        if not torch.cuda.is_available():
            raise L.PPASRB200Error("ppasr_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.torch = torch
        self.cfg = cfg
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.lib = L.load()
        c = Config()
        kind = getattr(cfg, "model_type", "conformer")
        squeeze = kind == "squeezeformer"
        c.reduce_idx = c.recover_idx = -1
        if kind == "deepspeech2":
            c.model_type = 2
            c.feat_dim = cfg.input_dim
            c.d_model = cfg.rnn_size
            c.n_layers = cfg.num_rnn_layers
            c.causal = int(cfg.streaming)
            c.use_gru = int(cfg.use_gru)
            c.vocab_size = cfg.vocab_size
            c.max_len = 5000
            self._create(c, weights)
            return
        c.model_type = 1 if squeeze else (3 if kind == "efficient_conformer" else 0)
        c.stride_layer_idx = -1
        if kind == "efficient_conformer":
            c.stride_layer_idx = -1 if cfg.stride_layer_idx is None else int(cfg.stride_layer_idx)
            c.group_layer_mask = sum(1 << int(i) for i in cfg.group_layer_idx)
            c.group_size = cfg.group_size
            c.stride_kernel = int(cfg.stride_kernel)
        c.feat_dim = cfg.input_dim
        c.d_model = cfg.encoder_dim if squeeze else cfg.output_size
        if squeeze:
            if cfg.output_size != cfg.encoder_dim:
                raise L.PPASRB200Error("squeezeformer final_proj (output_size != encoder_dim) is not supported")
            c.reduce_idx = -1 if cfg.reduce_idx is None else int(cfg.reduce_idx)
            c.recover_idx = -1 if cfg.recover_idx is None else int(cfg.recover_idx)
            c.time_reduce_kernel = 1 if cfg.time_reduction_layer_type == "stream" else 5
        c.n_heads = cfg.attention_heads
        c.ffn_dim = cfg.linear_units
        c.n_layers = cfg.num_blocks
        c.conv_kernel = cfg.cnn_module_kernel
        c.causal = int(cfg.causal)
        c.conv_norm = 0 if cfg.cnn_module_norm == "layer_norm" else 1
        c.vocab_size = cfg.vocab_size
        c.max_len = cfg.max_len
        self._create(c, weights)

    def _create(self, c, weights):
        self._ctx = ctypes.c_void_p()
        L.check(self.lib.ppasr_b200_create(ctypes.byref(c), ctypes.byref(self._ctx)))
        for name, arr in weights.items():
            # a trained checkpoint also holds the attention decoder of the *former models (`decoder.*`, 166 tensors that
            # CTC inference never reads; in DeepSpeech2 `decoder` IS the CTC head) and may hold 0-d bookkeeping entries
            if (c.model_type != 2 and name.startswith("decoder.")) or np.ndim(arr) == 0 or np.ndim(arr) > 4:
                continue
            a = np.ascontiguousarray(arr, dtype=np.float32)
            shape = (ctypes.c_int64 * a.ndim)(*a.shape)
            L.check(self.lib.ppasr_b200_load_tensor(self._ctx, name.encode(), a.ctypes.data_as(ctypes.c_void_p),
                                                    a.ndim, shape))
        L.check(self.lib.ppasr_b200_finalize(self._ctx))
        self.B = 0
        self.Tp = 0

    def close(self):
        if self._ctx:
            self.lib.ppasr_b200_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------------------------------------
    def encode(self, feats, lens=None, stream=None):
        """feats: float32 [B,T,F], torch CUDA tensor or host numpy/torch tensor. lens: int64 [B] (host)."""
        torch = self.torch
        if isinstance(feats, np.ndarray):
            feats = np.ascontiguousarray(feats, dtype=np.float32)
            B, T, _ = feats.shape
            ptr, on_dev = ctypes.c_void_p(feats.ctypes.data), 0
        else:
            feats = feats.contiguous()
            assert feats.dtype == torch.float32
            B, T, _ = feats.shape
            ptr, on_dev = ctypes.c_void_p(feats.data_ptr()), int(feats.is_cuda)
        lens_p = None
        if lens is not None:
            lens = np.ascontiguousarray(np.asarray(lens), dtype=np.int64)
            lens_p = ctypes.c_void_p(lens.ctypes.data)
        L.check(self.lib.ppasr_b200_encode(self._ctx, ptr, on_dev, lens_p, B, T, L.stream_ptr(stream)))
        self.B, self.Tp = B, int(self.lib.ppasr_b200_out_frames(self._ctx, T))
        return self

    # ---- ragged sessions (continuous batching of independent streams) -----------------------------------------------
    def sessions_init(self, max_sessions):
        L.check(self.lib.ppasr_b200_sessions_init(self._ctx, int(max_sessions)))

    def sessions_reset(self, slot):
        L.check(self.lib.ppasr_b200_sessions_reset(self._ctx, int(slot)))

    def sessions_step(self, feats, slots, required_cache_size=-1, stream=None):
        """feats float32 [n, t, F] (host NumPy or CUDA tensor), slots: the n distinct cache slots these chunks belong to."""
        torch = self.torch
        slots = np.ascontiguousarray(np.asarray(slots), dtype=np.int32)
        if isinstance(feats, np.ndarray):
            feats = np.ascontiguousarray(feats, dtype=np.float32)
            n, T, _ = feats.shape
            ptr, on_dev = ctypes.c_void_p(feats.ctypes.data), 0
        else:
            feats = feats.contiguous()
            n, T, _ = feats.shape
            ptr, on_dev = ctypes.c_void_p(feats.data_ptr()), int(feats.is_cuda)
        assert slots.shape == (n,)
        L.check(self.lib.ppasr_b200_sessions_step(self._ctx, ptr, on_dev, ctypes.c_void_p(slots.ctypes.data), n, T,
                                                  int(required_cache_size), L.stream_ptr(stream)))
        self.B, self.Tp = n, int(self.lib.ppasr_b200_out_frames(self._ctx, T))
        return self

    def ds2_states(self):
        """DeepSpeech2: (h, c) numpy [num_rnn_layers * num_directions, B, rnn_size] after the last encode_chunk."""
        cfg = self.cfg
        n = cfg.num_rnn_layers * cfg.num_directions
        h = np.empty((n, self.B, cfg.rnn_size), dtype=np.float32)
        c = np.empty_like(h)
        L.check(self.lib.ppasr_b200_ds2_states(self._ctx, h.ctypes.data_as(ctypes.c_void_p), c.ctypes.data_as(ctypes.c_void_p),
                                               0, None))
        return h, c

    # ---- chunk streaming -------------------------------------------------------------------------
    def stream_reset(self, B=1):
        L.check(self.lib.ppasr_b200_stream_reset(self._ctx, int(B)))

    def encode_chunk(self, feats, required_cache_size=-1, stream=None):
        """feats: float32 [B,t,F] (host numpy / torch, or CUDA tensor). Appends out_frames(t) frames to the stream."""
        torch = self.torch
        if isinstance(feats, np.ndarray):
            feats = np.ascontiguousarray(feats, dtype=np.float32)
            B, T, _ = feats.shape
            ptr, on_dev = ctypes.c_void_p(feats.ctypes.data), 0
        else:
            feats = feats.contiguous()
            B, T, _ = feats.shape
            ptr, on_dev = ctypes.c_void_p(feats.data_ptr()), int(feats.is_cuda)
        L.check(self.lib.ppasr_b200_encode_chunk(self._ctx, ptr, on_dev, B, T, int(required_cache_size),
                                                 L.stream_ptr(stream)))
        self.B, self.Tp = B, int(self.lib.ppasr_b200_out_frames(self._ctx, T))
        return self

    def stream_info(self):
        off, ct = ctypes.c_int32(), ctypes.c_int32()
        L.check(self.lib.ppasr_b200_stream_info(self._ctx, ctypes.byref(off), ctypes.byref(ct)))
        return off.value, ct.value

    def stream_export(self):
        """(att_cache [L,H,t,128], cnn_cache [L,1,D,K-1]) fp32 host copies in the reference layouts (stream 0)."""
        cfg = self.cfg
        _, ct = self.stream_info()
        att = np.zeros((cfg.num_blocks, cfg.attention_heads, ct, 2 * (cfg.output_size // cfg.attention_heads)),
                       dtype=np.float32)
        cnn = np.zeros((cfg.num_blocks, 1, cfg.output_size, cfg.cnn_module_kernel - 1), dtype=np.float32)
        L.check(self.lib.ppasr_b200_stream_export(self._ctx, ctypes.c_void_p(att.ctypes.data),
                                                  ctypes.c_void_p(cnn.ctypes.data), 0, L.stream_ptr()))
        return att, cnn

    def ctc_probs(self, to_host=False, stream=None):
        torch = self.torch
        V = self.cfg.vocab_size
        if to_host:
            out = np.empty((self.B, self.Tp, V), dtype=np.float32)
            L.check(self.lib.ppasr_b200_ctc_probs(self._ctx, ctypes.c_void_p(out.ctypes.data), 0, L.stream_ptr(stream)))
            return out
        out = torch.empty((self.B, self.Tp, V), dtype=torch.float32, device=self.device)
        L.check(self.lib.ppasr_b200_ctc_probs(self._ctx, L.ptr(out), 1, L.stream_ptr(stream)))
        return out

    def ctc_logits(self, stream=None):
        torch = self.torch
        out = torch.empty((self.B, self.Tp, self.cfg.vocab_size), dtype=torch.float32, device=self.device)
        L.check(self.lib.ppasr_b200_ctc_logits(self._ctx, L.ptr(out), 1, L.stream_ptr(stream)))
        return out

    def ctc_greedy(self, to_host=True, trim_to_lens=False, blank_id=0, with_frames=False, stream=None):
        """Fused CTC head + greedy decode. Returns (ids [B,T'], out_lens [B], scores [B][, frame_ids, frame_probs])."""
        torch = self.torch
        B, Tp = self.B, self.Tp
        if to_host:
            ids = np.empty((B, Tp), dtype=np.int32)
            ol = np.empty((B,), dtype=np.int32)
            sc = np.empty((B,), dtype=np.float32)
            fi = np.empty((B, Tp), dtype=np.int32) if with_frames else None
            fp = np.empty((B, Tp), dtype=np.float32) if with_frames else None
            p = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else None
            L.check(self.lib.ppasr_b200_ctc_greedy(self._ctx, p(ids), p(ol), p(sc), p(fi), p(fp), 0,
                                                   int(trim_to_lens), blank_id, L.stream_ptr(stream)))
        else:
            ids = torch.empty((B, Tp), dtype=torch.int32, device=self.device)
            ol = torch.empty((B,), dtype=torch.int32, device=self.device)
            sc = torch.empty((B,), dtype=torch.float32, device=self.device)
            fi = torch.empty((B, Tp), dtype=torch.int32, device=self.device) if with_frames else None
            fp = torch.empty((B, Tp), dtype=torch.float32, device=self.device) if with_frames else None
            L.check(self.lib.ppasr_b200_ctc_greedy(self._ctx, L.ptr(ids), L.ptr(ol), L.ptr(sc), L.ptr(fi), L.ptr(fp), 1,
                                                   int(trim_to_lens), blank_id, L.stream_ptr(stream)))
        if with_frames:
            return ids, ol, sc, fi, fp
        return ids, ol, sc

    def ctc_greedy_into(self, ids, out_lens, scores, trim_to_lens=False, blank_id=0, stream=None):
        """Fused CTC head + greedy decode into caller-owned CUDA tensors (ids int32 [B,T'], out_lens int32 [B], scores fp32 [B]):
        no allocation, no host synchronisation -- the form a CUDA graph capture needs."""
        L.check(self.lib.ppasr_b200_ctc_greedy(self._ctx, L.ptr(ids), L.ptr(out_lens), L.ptr(scores), None, None, 1,
                                               int(trim_to_lens), blank_id, L.stream_ptr(stream)))

    # ---- CUDA graph of the steady-state step (include/ppasr_b200.h "CUDA graph") -----------------------------------------
    def graph_begin(self, stream):
        L.check(self.lib.ppasr_b200_graph_begin(self._ctx, L.stream_ptr(stream)))

    def graph_end(self, stream):
        L.check(self.lib.ppasr_b200_graph_end(self._ctx, L.stream_ptr(stream)))
        return int(self.lib.ppasr_b200_graph_kernels(self._ctx))

    def graph_launch(self, stream):
        L.check(self.lib.ppasr_b200_graph_launch(self._ctx, L.stream_ptr(stream)))

    def set_option(self, name, value):
        L.check(self.lib.ppasr_b200_set_option(self._ctx, name.encode(), int(value)))

    def profile_enable(self, on=True, only=None):
        """on: time every kernel class; only='fused_ffn': event pairs around that class alone (the rest runs undisturbed)."""
        mode = int(bool(on))
        if on and only is not None:
            n = self.lib.ppasr_b200_profile_num_classes()
            names = [self.lib.ppasr_b200_profile_class_name(i).decode() for i in range(n)]
            mode = 2 + names.index(only)
        L.check(self.lib.ppasr_b200_profile_enable(self._ctx, mode))

    def profile_read(self):
        """Returns {class_name: (launches, total_ms)} for the launches since profile_enable(True)."""
        n = self.lib.ppasr_b200_profile_num_classes()
        counts = (ctypes.c_int32 * n)()
        ms = (ctypes.c_float * n)()
        L.check(self.lib.ppasr_b200_profile_read(self._ctx, counts, ms))
        return {self.lib.ppasr_b200_profile_class_name(i).decode(): (int(counts[i]), float(ms[i]))
                for i in range(n) if counts[i] > 0}

    def debug_phase(self):
        """conv1 output as the stride-phase images: bf16 [4, B, Th, FH, D] (ppasr_b200_debug_copy_phase)."""
        torch = self.torch
        dims = (ctypes.c_int32 * 4)()
        L.check(self.lib.ppasr_b200_debug_copy_phase(self._ctx, None, dims, L.stream_ptr()))
        B, Th, FH, D = (int(v) for v in dims)
        out = torch.empty((4, B, Th, FH, D), dtype=torch.bfloat16, device=self.device)
        L.check(self.lib.ppasr_b200_debug_copy_phase(self._ctx, L.ptr(out), dims, L.stream_ptr()))
        return out

    def debug_x(self):
        torch = self.torch
        out = torch.empty((self.B * self.Tp, self.cfg.output_size), dtype=torch.float32, device=self.device)
        L.check(self.lib.ppasr_b200_debug_copy_x(self._ctx, L.ptr(out), L.stream_ptr()))
        return out
