"""ctypes binding of include/coati_gpu.h (no torch types, no compute on the Python side)."""
from __future__ import annotations

import ctypes as C
import time
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_fp = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)


class CoatiGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"coati_gpu error {code}: {msg}")
        self.code = code


def library_path() -> str:
    if os.environ.get("COATI_GPU_LIB"):   # A/B builds of the same library (tools/gpu experiments)
        return os.environ["COATI_GPU_LIB"]
    return os.path.join(_HERE, "libcoati_gpu.so")


def load_library() -> C.CDLL:
    """Load libcoati_gpu.so.  There is no fallback: a missing library is an error."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the CUDA extension is the product; there is no CPU path)")
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.coati_gpu_init.argtypes = [C.c_int, C.POINTER(vp)]
    lib.coati_gpu_shutdown.argtypes = [vp]
    lib.coati_gpu_shutdown.restype = None
    lib.coati_gpu_strerror.argtypes = [C.c_int]
    lib.coati_gpu_strerror.restype = C.c_char_p
    lib.coati_gpu_last_cuda_error.argtypes = [vp]
    lib.coati_gpu_last_cuda_error.restype = C.c_char_p
    lib.coati_gpu_stream.argtypes = [vp]
    lib.coati_gpu_stream.restype = vp
    lib.coati_gpu_launch_count.argtypes = [vp]
    lib.coati_gpu_launch_count.restype = C.c_uint64
    lib.coati_gpu_transfer_bytes.argtypes = [vp, _u64p, _u64p]
    lib.coati_gpu_transfer_bytes.restype = None
    lib.coati_gpu_device_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.coati_gpu_set_model.argtypes = [vp, _fp, C.c_float, C.c_float, C.c_uint32]
    lib.coati_gpu_set_models.argtypes = [vp, C.c_uint32, _fp, C.c_float, C.c_float, C.c_uint32]
    lib.coati_gpu_viterbi_batch_models.argtypes = [vp, C.c_size_t, vp, _u64p, vp, _u64p, vp, vp,
                                                   C.POINTER(C.c_uint32), vp, vp, _u64p, _fp, _i32p]
    lib.coati_gpu_viterbi.argtypes = [vp, _u8p, C.c_size_t, _u8p, C.c_size_t, C.c_char_p, C.c_char_p,
                                      C.c_char_p, C.c_char_p, C.POINTER(C.c_size_t), _fp]
    lib.coati_gpu_viterbi_batch.argtypes = [vp, C.c_size_t, vp, _u64p, vp, _u64p, vp, vp, vp, vp,
                                            _u64p, _fp, _i32p]
    lib.coati_gpu_alignpair_batch.argtypes = [vp, C.c_size_t, vp, _u64p, vp, _u64p, vp, vp, _u64p, _fp, _i32p]
    lib.coati_gpu_multi_alignpair_batch.argtypes = [C.POINTER(vp), C.c_int, C.c_size_t, vp, _u64p, vp, _u64p, vp, vp,
                                                    _u64p, _fp, _i32p]
    lib.coati_gpu_plan_shards.argtypes = [C.c_size_t, _u64p, _u64p, C.c_uint32, C.c_size_t, _u64p, _u64p,
                                          C.POINTER(C.c_uint32)]
    lib.coati_gpu_plan_shards.restype = C.c_size_t
    lib.coati_gpu_alignpair_batch_ranges.argtypes = [vp, C.c_size_t, vp, _u64p, vp, _u64p, vp, vp, _u64p, _fp, _i32p,
                                                     C.c_size_t, _u64p, _u64p]
    lib.coati_gpu_host_alloc.argtypes = [C.c_size_t]
    lib.coati_gpu_host_alloc.restype = vp
    lib.coati_gpu_host_free.argtypes = [vp]
    lib.coati_gpu_host_free.restype = None
    lib.coati_gpu_host_register.argtypes = [vp, C.c_size_t]
    lib.coati_gpu_host_unregister.argtypes = [vp]
    lib.coati_gpu_batch_create.argtypes = [vp, C.c_size_t, _u64p, _u64p, C.POINTER(vp)]
    lib.coati_gpu_batch_upload.argtypes = [vp, vp, vp, vp, vp]
    lib.coati_gpu_batch_run.argtypes = [vp]
    lib.coati_gpu_batch_download.argtypes = [vp, vp, vp, _u64p, _fp, _i32p]
    lib.coati_gpu_batch_stats.argtypes = [vp, _u64p, _u64p, _u64p, _u64p]
    lib.coati_gpu_batch_timing.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                           C.POINTER(C.c_double), _u64p]
    lib.coati_gpu_batch_device_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), _u64p,
                                                   C.POINTER(vp), _u64p]
    lib.coati_gpu_batch_destroy.argtypes = [vp]
    lib.coati_gpu_batch_destroy.restype = None
    # host layer (no GPU): table builder, encoding, seeding, re-scoring
    lib.coati_host_marginal_table.argtypes = [C.c_int, C.c_float, C.c_float, _fp, C.c_int, C.c_int, _fp]
    lib.coati_host_marginal_table_gtr.argtypes = [C.c_float, C.c_float, _fp, _fp, _fp]
    lib.coati_host_mg94_p.argtypes = [C.c_float, C.c_float, _fp, _fp, _fp]
    lib.coati_host_gtr_q.argtypes = [_fp, _fp, _fp]
    lib.coati_host_encode.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, _u8p, _u8p]
    lib.coati_host_seed.argtypes = [C.POINTER(C.c_char_p), C.c_size_t, _u64p]
    lib.coati_host_seed.restype = None
    lib.coati_host_alignment_score.argtypes = [C.c_char_p, C.c_char_p, _fp, C.c_float, C.c_float, C.c_size_t, _fp]
    lib.coati_host_json_number.argtypes = [C.c_float, C.c_char_p, C.c_size_t]
    lib.coati_host_json_number.restype = None
    lib.coati_gpu_forward.argtypes = [vp, _u8p, C.c_size_t, _u8p, C.c_size_t, C.POINTER(vp)]
    lib.coati_gpu_forward_terminal.argtypes = [vp, _fp, _fp]
    lib.coati_gpu_sampleback.argtypes = [vp, C.c_char_p, C.c_char_p, _u64p, C.c_size_t, vp, vp,
                                         C.POINTER(C.c_size_t), _fp, _fp]
    lib.coati_gpu_forward_batch.argtypes = [vp, C.c_size_t, vp, _u64p, vp, _u64p, C.POINTER(vp)]
    lib.coati_gpu_forward_batch_terminal.argtypes = [vp, _fp, _fp, _fp]
    lib.coati_gpu_sampleback_batch.argtypes = [vp, vp, vp, _u64p, C.c_size_t, vp, vp, _u64p, _fp, _fp]
    lib.coati_gpu_forward_free.argtypes = [vp]
    lib.coati_gpu_forward_free.restype = None
    lib.coati_gpu_forward_matrices.argtypes = [vp, _fp, _fp, _fp]
    lib.coati_gpu_libm_eval.argtypes = [vp, C.c_int, _fp, _fp, C.c_size_t]
    lib.coati_gpu_viterbi_directions.argtypes = [vp, _u8p, C.c_size_t, _u8p, C.c_size_t, _u8p, _fp]
    _LIB = lib
    return lib


def _vp(arr: np.ndarray):
    return C.c_void_p(arr.ctypes.data)


class PackedPairs:
    """CSR pack of a batch (the layout of coati_gpu_viterbi_batch)."""

    def __init__(self, a_list, b_list, anc_list, des_list):
        n = len(a_list)
        self.n = n
        la = np.fromiter((len(x) for x in a_list), dtype=np.uint64, count=n)
        lb = np.fromiter((len(x) for x in b_list), dtype=np.uint64, count=n)
        self.a_off = np.zeros(n + 1, dtype=np.uint64)
        self.b_off = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(la, out=self.a_off[1:])
        np.cumsum(lb, out=self.b_off[1:])
        cat = lambda xs, dt: (np.concatenate([np.asarray(x, dtype=dt) for x in xs])  # noqa: E731
                              if n and sum(len(x) for x in xs) else np.zeros(0, dtype=dt))
        self.a_all = np.ascontiguousarray(cat(a_list, np.uint8))
        self.b_all = np.ascontiguousarray(cat(b_list, np.uint8))
        self.anc_all = np.frombuffer("".join(anc_list).encode("latin-1"), dtype=np.uint8).copy()
        self.des_all = np.frombuffer("".join(des_list).encode("latin-1"), dtype=np.uint8).copy()
        assert len(self.anc_all) == len(self.a_all) and len(self.des_all) == len(self.b_all)
        self.out_off = self.a_off[:-1] + self.b_off[:-1] + np.arange(n, dtype=np.uint64)
        self.out_total = int(self.a_off[-1] + self.b_off[-1]) + n

    def cells(self) -> int:
        la = np.diff(self.a_off).astype(np.float64)
        lb = np.diff(self.b_off).astype(np.float64)
        return int((la * lb).sum())


class Batch:
    def __init__(self, ctx: "Context", a_off: np.ndarray, b_off: np.ndarray):
        self.ctx = ctx
        self.lib = ctx.lib
        self.n = len(a_off) - 1
        self.h = C.c_void_p()
        self._a_off = np.ascontiguousarray(a_off, dtype=np.uint64)
        self._b_off = np.ascontiguousarray(b_off, dtype=np.uint64)
        ctx._check(self.lib.coati_gpu_batch_create(ctx.h, self.n, self._a_off.ctypes.data_as(_u64p),
                                                   self._b_off.ctypes.data_as(_u64p), C.byref(self.h)))

    def upload(self, a_all, b_all, anc_all, des_all):
        self.ctx._check(self.lib.coati_gpu_batch_upload(self.h, _vp(a_all), _vp(b_all), _vp(anc_all),
                                                        _vp(des_all)))

    def run(self):
        self.ctx._check(self.lib.coati_gpu_batch_run(self.h))

    def download(self, out_a, out_b, out_len, score, status):
        self.ctx._check(self.lib.coati_gpu_batch_download(
            self.h, _vp(out_a), _vp(out_b), out_len.ctypes.data_as(_u64p), score.ctypes.data_as(_fp),
            status.ctypes.data_as(_i32p)))

    def stats(self):
        v = [C.c_uint64(0) for _ in range(4)]
        self.ctx._check(self.lib.coati_gpu_batch_stats(self.h, *[C.byref(x) for x in v]))
        return dict(cells=v[0].value, dir_bytes=v[1].value, launches=v[2].value, chunks=v[3].value)

    def timing(self):
        f, t, c = C.c_double(0), C.c_double(0), C.c_double(0)
        n = C.c_uint64(0)
        self.ctx._check(self.lib.coati_gpu_batch_timing(self.h, C.byref(f), C.byref(t), C.byref(c),
                                                        C.byref(n)))
        return dict(fill_ms=f.value, traceback_ms=t.value, compact_ms=c.value, fill_launches=n.value)

    def device_buffers(self):
        oa, ob, rs = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nb, rb = C.c_uint64(0), C.c_uint64(0)
        self.ctx._check(self.lib.coati_gpu_batch_device_buffers(self.h, C.byref(oa), C.byref(ob),
                                                                C.byref(nb), C.byref(rs), C.byref(rb)))
        return dict(out_a=oa.value, out_b=ob.value, out_bytes=nb.value, results=rs.value,
                    result_bytes=rb.value)

    def destroy(self):
        if self.h:
            self.lib.coati_gpu_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Forward:
    """Opaque Forward work object (coati_gpu_forward_t)."""

    def __init__(self, ctx: "Context", a, b):
        self.ctx, self.lib = ctx, ctx.lib
        self.a = np.ascontiguousarray(a, dtype=np.uint8)
        self.b = np.ascontiguousarray(b, dtype=np.uint8)
        self.h = C.c_void_p()
        ctx._check(self.lib.coati_gpu_forward(ctx.h, self.a.ctypes.data_as(_u8p), len(self.a),
                                              self.b.ctypes.data_as(_u8p), len(self.b), C.byref(self.h)))

    def terminal(self):
        term = np.zeros(3, dtype=np.float32)
        ms = C.c_float(0)
        self.ctx._check(self.lib.coati_gpu_forward_terminal(self.h, term.ctypes.data_as(_fp), C.byref(ms)))
        return term, ms.value

    def matrices(self):
        shape = (len(self.a) + 1, len(self.b) + 1)
        m, d, i = (np.empty(shape, dtype=np.float32) for _ in range(3))
        self.ctx._check(self.lib.coati_gpu_forward_matrices(self.h, m.ctypes.data_as(_fp), d.ctypes.data_as(_fp),
                                                            i.ctypes.data_as(_fp)))
        return m, d, i

    def sampleback(self, anc: str, des: str, state, n: int):
        """Returns (list[(row_a, row_b)], float32 scores, new_state, sample_ms)."""
        stride = len(self.a) + len(self.b) + 1
        oa = np.zeros(n * stride + 1, dtype=np.uint8)
        ob = np.zeros(n * stride + 1, dtype=np.uint8)
        ol = (C.c_size_t * max(n, 1))()
        sc = np.zeros(max(n, 1), dtype=np.float32)
        st = (C.c_uint64 * 2)(int(state[0]), int(state[1]))
        ms = C.c_float(0)
        anc_b, des_b = anc.encode("latin-1"), des.encode("latin-1")
        t0 = time.perf_counter()
        self.ctx._check(self.lib.coati_gpu_sampleback(self.h, anc_b, des_b, st, n,
                                                      _vp(oa), _vp(ob), ol, sc.ctypes.data_as(_fp), C.byref(ms)))
        self.last_call_s = time.perf_counter() - t0  # the C-ABI call alone (kernels + D2H), no Python decoding
        rows = [(oa[s * stride:s * stride + ol[s]].tobytes().decode("latin-1"),
                 ob[s * stride:s * stride + ol[s]].tobytes().decode("latin-1")) for s in range(n)]
        return rows, sc[:n], np.array([st[0], st[1]], dtype=np.uint64), ms.value

    def free(self):
        if self.h:
            self.lib.coati_gpu_forward_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ForwardBatch:
    """Forward matrices of a batch of pairs (coati_gpu_forward_batch) + per-pair seeded sampling."""

    def __init__(self, ctx: "Context", pack: "PackedPairs"):
        self.ctx, self.lib, self.pack = ctx, ctx.lib, pack
        self.h = C.c_void_p()
        ctx._check(self.lib.coati_gpu_forward_batch(ctx.h, pack.n, _vp(pack.a_all), pack.a_off.ctypes.data_as(_u64p),
                                                    _vp(pack.b_all), pack.b_off.ctypes.data_as(_u64p), C.byref(self.h)))

    def terminal(self):
        """(term[npairs, 3], loglik[npairs], fill_ms)"""
        n = self.pack.n
        term = np.zeros((n, 3), dtype=np.float32)
        ll = np.zeros(n, dtype=np.float32)
        ms = C.c_float(0)
        self.ctx._check(self.lib.coati_gpu_forward_batch_terminal(self.h, term.ctypes.data_as(_fp),
                                                                  ll.ctypes.data_as(_fp), C.byref(ms)))
        return term, ll, ms.value

    def sampleback(self, states, n: int):
        """states: uint64[npairs, 2].  Returns (rows[p][s] = (row_a, row_b), scores[npairs, n], new states, ms)."""
        pk = self.pack
        npairs = pk.n
        total = n * (int(pk.a_off[-1]) + int(pk.b_off[-1]) + npairs)
        oa = np.zeros(total + 1, dtype=np.uint8)
        ob = np.zeros(total + 1, dtype=np.uint8)
        ol = np.zeros(max(1, npairs * n), dtype=np.uint64)
        sc = np.zeros(max(1, npairs * n), dtype=np.float32)
        st = np.ascontiguousarray(states, dtype=np.uint64).reshape(npairs, 2).copy()
        ms = C.c_float(0)
        self.ctx._check(self.lib.coati_gpu_sampleback_batch(self.h, _vp(pk.anc_all), _vp(pk.des_all),
                                                            st.ctypes.data_as(_u64p), n, _vp(oa), _vp(ob),
                                                            ol.ctypes.data_as(_u64p), sc.ctypes.data_as(_fp), C.byref(ms)))
        rows = []
        for p in range(npairs):
            base = n * (int(pk.a_off[p]) + int(pk.b_off[p]) + p)
            stride = int(pk.a_off[p + 1] - pk.a_off[p]) + int(pk.b_off[p + 1] - pk.b_off[p]) + 1
            rows.append([(oa[base + s * stride: base + s * stride + int(ol[p * n + s])].tobytes().decode("latin-1"),
                          ob[base + s * stride: base + s * stride + int(ol[p * n + s])].tobytes().decode("latin-1"))
                         for s in range(n)])
        return rows, sc[:npairs * n].reshape(npairs, n), st, ms.value

    def free(self):
        if self.h:
            self.lib.coati_gpu_forward_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One GPU context (coati_gpu_ctx).  Raises CoatiGpuError on any failure -- never falls back."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.coati_gpu_init(device, C.byref(self.h))
        if rc != 0:
            raise CoatiGpuError(rc, self.lib.coati_gpu_strerror(rc).decode())
        self.k = None

    def _check(self, rc: int):
        if rc != 0:
            extra = self.lib.coati_gpu_last_cuda_error(self.h).decode() if self.h else ""
            raise CoatiGpuError(rc, self.lib.coati_gpu_strerror(rc).decode() + (" [" + extra + "]" if extra else ""))

    def close(self):
        if self.h:
            self.lib.coati_gpu_shutdown(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        return int(self.lib.coati_gpu_stream(self.h) or 0)

    @property
    def launches(self) -> int:
        return int(self.lib.coati_gpu_launch_count(self.h))

    @property
    def transfer_bytes(self):
        """(host->device, device->host) bytes moved by the batch calls of this context since creation"""
        h2d, d2h = C.c_uint64(0), C.c_uint64(0)
        self.lib.coati_gpu_transfer_bytes(self.h, C.byref(h2d), C.byref(d2h))
        return int(h2d.value), int(d2h.value)

    def device_info(self):
        sm, khz = C.c_int(0), C.c_int(0)
        fr, tot = C.c_size_t(0), C.c_size_t(0)
        self._check(self.lib.coati_gpu_device_info(self.h, C.byref(sm), C.byref(khz), C.byref(fr), C.byref(tot)))
        return dict(sm_count=sm.value, clock_khz=khz.value, free_bytes=fr.value, total_bytes=tot.value)

    def set_model(self, table, gap_open=0.001, gap_extend=1.0 - 1.0 / 6.0, gap_len=1):
        t = np.ascontiguousarray(table, dtype=np.float32)
        assert t.shape == (183, 15)
        self._check(self.lib.coati_gpu_set_model(self.h, t.ctypes.data_as(_fp), np.float32(gap_open),
                                                 np.float32(gap_extend), int(gap_len)))
        self.k = int(gap_len)

    def viterbi(self, a, b, anc: str, des: str):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        b = np.ascontiguousarray(b, dtype=np.uint8)
        n = len(a) + len(b) + 1
        oa, ob = C.create_string_buffer(n), C.create_string_buffer(n)
        ol, sc = C.c_size_t(0), C.c_float(0)
        self._check(self.lib.coati_gpu_viterbi(self.h, a.ctypes.data_as(_u8p), len(a), b.ctypes.data_as(_u8p),
                                               len(b), anc.encode("latin-1"), des.encode("latin-1"), oa, ob,
                                               C.byref(ol), C.byref(sc)))
        return oa.raw[:ol.value].decode("latin-1"), ob.raw[:ol.value].decode("latin-1"), np.float32(sc.value)

    def set_models(self, tables, gap_open=0.001, gap_extend=1.0 - 1.0 / 6.0, gap_len=1):
        t = np.ascontiguousarray(tables, dtype=np.float32)
        assert t.ndim == 3 and t.shape[1:] == (183, 15)
        self._check(self.lib.coati_gpu_set_models(self.h, t.shape[0], t.ctypes.data_as(_fp), np.float32(gap_open),
                                                  np.float32(gap_extend), int(gap_len)))
        self.k = int(gap_len)

    def viterbi_batch(self, pack: PackedPairs, model_idx=None):
        """Returns (rows_a, rows_b, scores float32[n], status int32[n]) in input order."""
        if model_idx is not None:
            return self._viterbi_batch_models(pack, model_idx)
        out_a = np.zeros(pack.out_total + 1, dtype=np.uint8)
        out_b = np.zeros(pack.out_total + 1, dtype=np.uint8)
        out_len = np.zeros(pack.n, dtype=np.uint64)
        score = np.zeros(pack.n, dtype=np.float32)
        status = np.zeros(pack.n, dtype=np.int32)
        self._check(self.lib.coati_gpu_viterbi_batch(
            self.h, pack.n, _vp(pack.a_all), pack.a_off.ctypes.data_as(_u64p), _vp(pack.b_all),
            pack.b_off.ctypes.data_as(_u64p), _vp(pack.anc_all), _vp(pack.des_all), _vp(out_a), _vp(out_b),
            out_len.ctypes.data_as(_u64p), score.ctypes.data_as(_fp), status.ctypes.data_as(_i32p)))
        rows_a, rows_b = [], []
        for p in range(pack.n):
            o, n = int(pack.out_off[p]), int(out_len[p])
            rows_a.append(out_a[o:o + n].tobytes().decode("latin-1"))
            rows_b.append(out_b[o:o + n].tobytes().decode("latin-1"))
        return rows_a, rows_b, score, status

    def _viterbi_batch_models(self, pack: PackedPairs, model_idx):
        m = np.ascontiguousarray(model_idx, dtype=np.uint32)
        out_a = np.zeros(pack.out_total + 1, dtype=np.uint8)
        out_b = np.zeros(pack.out_total + 1, dtype=np.uint8)
        out_len = np.zeros(pack.n, dtype=np.uint64)
        score = np.zeros(pack.n, dtype=np.float32)
        status = np.zeros(pack.n, dtype=np.int32)
        self._check(self.lib.coati_gpu_viterbi_batch_models(
            self.h, pack.n, _vp(pack.a_all), pack.a_off.ctypes.data_as(_u64p), _vp(pack.b_all),
            pack.b_off.ctypes.data_as(_u64p), _vp(pack.anc_all), _vp(pack.des_all),
            m.ctypes.data_as(C.POINTER(C.c_uint32)), _vp(out_a), _vp(out_b), out_len.ctypes.data_as(_u64p),
            score.ctypes.data_as(_fp), status.ctypes.data_as(_i32p)))
        rows_a, rows_b = [], []
        for p in range(pack.n):
            o, n = int(pack.out_off[p]), int(out_len[p])
            rows_a.append(out_a[o:o + n].tobytes().decode("latin-1"))
            rows_b.append(out_b[o:o + n].tobytes().decode("latin-1"))
        return rows_a, rows_b, score, status

    def forward_batch(self, pack: "PackedPairs") -> "ForwardBatch":
        return ForwardBatch(self, pack)

    def forward(self, a, b) -> "Forward":
        return Forward(self, a, b)

    def libm_eval(self, op: int, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self._check(self.lib.coati_gpu_libm_eval(self.h, op, x.ctypes.data_as(_fp), out.ctypes.data_as(_fp), x.size))
        return out

    def alignpair_batch(self, ancs, dess):
        """marg_alignment semantics for a list of raw pairs.  Returns (rows_a, rows_b, scores, status)."""
        n = len(ancs)
        la = np.fromiter((len(x) for x in ancs), dtype=np.uint64, count=n)
        lb = np.fromiter((len(x) for x in dess), dtype=np.uint64, count=n)
        a_off = np.zeros(n + 1, dtype=np.uint64)
        b_off = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(la, out=a_off[1:])
        np.cumsum(lb, out=b_off[1:])
        anc_all = np.frombuffer(("".join(ancs) + "\0").encode("latin-1"), dtype=np.uint8).copy()
        des_all = np.frombuffer(("".join(dess) + "\0").encode("latin-1"), dtype=np.uint8).copy()
        total = int(a_off[-1] + b_off[-1]) + n
        out_a = np.zeros(total + 1, dtype=np.uint8)
        out_b = np.zeros(total + 1, dtype=np.uint8)
        out_len = np.zeros(n, dtype=np.uint64)
        score = np.zeros(n, dtype=np.float32)
        status = np.zeros(n, dtype=np.int32)
        self._check(self.lib.coati_gpu_alignpair_batch(
            self.h, n, _vp(anc_all), a_off.ctypes.data_as(_u64p), _vp(des_all), b_off.ctypes.data_as(_u64p),
            _vp(out_a), _vp(out_b), out_len.ctypes.data_as(_u64p), score.ctypes.data_as(_fp),
            status.ctypes.data_as(_i32p)))
        rows_a, rows_b = [], []
        for p in range(n):
            o, ln = int(a_off[p] + b_off[p]) + p, int(out_len[p])
            rows_a.append(out_a[o:o + ln].tobytes().decode("latin-1"))
            rows_b.append(out_b[o:o + ln].tobytes().decode("latin-1"))
        return rows_a, rows_b, score, status

    def batch(self, a_off, b_off) -> Batch:
        return Batch(self, a_off, b_off)

    def directions(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        b = np.ascontiguousarray(b, dtype=np.uint8)
        d = np.zeros((len(a), len(b)), dtype=np.uint8)
        score = C.c_float(0)
        self._check(self.lib.coati_gpu_viterbi_directions(self.h, a.ctypes.data_as(_u8p), len(a),
                                                          b.ctypes.data_as(_u8p), len(b),
                                                          d.ctypes.data_as(_u8p), C.byref(score)))
        return d, np.float32(score.value)


def plan_shards(a_off, b_off, n_shards: int):
    """coati_gpu_plan_shards: contiguous chunks of one CSR batch, heaviest first, given to shards by greedy
    longest-processing-time.  Returns (first, last, shard) arrays, one entry per chunk."""
    lib = load_library()
    a_off = np.ascontiguousarray(a_off, dtype=np.uint64)
    b_off = np.ascontiguousarray(b_off, dtype=np.uint64)
    npairs = len(a_off) - 1
    cap = 2 * (npairs // 8192 + n_shards + 2)
    first = np.zeros(cap, dtype=np.uint64)
    last = np.zeros(cap, dtype=np.uint64)
    shard = np.zeros(cap, dtype=np.uint32)
    n = lib.coati_gpu_plan_shards(npairs, a_off.ctypes.data_as(_u64p), b_off.ctypes.data_as(_u64p), n_shards, cap,
                                  first.ctypes.data_as(_u64p), last.ctypes.data_as(_u64p),
                                  shard.ctypes.data_as(C.POINTER(C.c_uint32)))
    if n == 0 and npairs:
        raise RuntimeError("coati_gpu_plan_shards failed")
    return first[:n].copy(), last[:n].copy(), shard[:n].copy()


def alignpair_batch_ranges(ctx: "Context", w, outs, first, last):
    """coati_gpu_alignpair_batch_ranges on the CSR batch `w` (dict of a_off, b_off, anc_all, des_all) into
    outs = (out_a, out_b, out_len, score, status)."""
    out_a, out_b, out_len, score, status = outs
    first = np.ascontiguousarray(first, dtype=np.uint64)
    last = np.ascontiguousarray(last, dtype=np.uint64)
    rows = os.environ.get("COATI_DIAG_NO_ROWS") != "1"   # diagnostics: skip the D2H of the rows
    ctx._check(ctx.lib.coati_gpu_alignpair_batch_ranges(
        ctx.h, len(w["a_off"]) - 1, _vp(w["anc_all"]), w["a_off"].ctypes.data_as(_u64p), _vp(w["des_all"]),
        w["b_off"].ctypes.data_as(_u64p), _vp(out_a) if rows else None, _vp(out_b) if rows else None,
        out_len.ctypes.data_as(_u64p),
        score.ctypes.data_as(_fp), status.ctypes.data_as(_i32p), len(first), first.ctypes.data_as(_u64p),
        last.ctypes.data_as(_u64p)))


def multi_alignpair_batch(ctxs, w, outs):
    """coati_gpu_multi_alignpair_batch: one CSR batch over several contexts (one per device)."""
    out_a, out_b, out_len, score, status = outs
    lib = ctxs[0].lib
    arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    ctxs[0]._check(lib.coati_gpu_multi_alignpair_batch(
        arr, len(ctxs), len(w["a_off"]) - 1, _vp(w["anc_all"]), w["a_off"].ctypes.data_as(_u64p), _vp(w["des_all"]),
        w["b_off"].ctypes.data_as(_u64p), _vp(out_a), _vp(out_b), out_len.ctypes.data_as(_u64p),
        score.ctypes.data_as(_fp), status.ctypes.data_as(_i32p)))


class PinnedArena:
    """uint8 numpy view of page-locked host memory from coati_gpu_host_alloc."""

    def __init__(self, nbytes: int):
        self.lib = load_library()
        self.nbytes = max(1, int(nbytes))
        self.ptr = self.lib.coati_gpu_host_alloc(self.nbytes)
        if not self.ptr:
            raise MemoryError("coati_gpu_host_alloc failed")
        self.array = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.coati_gpu_host_free(self.ptr)
            self.ptr = None


# ---- host layer helpers (C++ table builder / sequence prep behind coati_host_* entry points) ---------
def host_marginal_table(model="mar-mg", br_len=0.0133, omega=0.2, pi=(0.308, 0.185, 0.199, 0.308),
                        amb="SUM", msub="SUM"):
    lib = load_library()
    out = np.zeros((183, 15), dtype=np.float32)
    p = np.asarray(pi, dtype=np.float32)
    rc = lib.coati_host_marginal_table({"mar-mg": 0, "mar-ecm": 1}[model], br_len, omega, p.ctypes.data_as(_fp),
                                       int(amb == "BEST"), int(msub == "MAX"), out.ctypes.data_as(_fp))
    if rc != 0:
        raise ValueError("set_subst failed")
    return out


def host_marginal_table_gtr(br_len=0.0133, omega=0.2, pi=(0.308, 0.185, 0.199, 0.308), sigma=(0,) * 6):
    """mar-mg table with the GTR rates wired through (alignment_t::use_sigma; upstream drops them)."""
    lib = load_library()
    out = np.zeros((183, 15), dtype=np.float32)
    p = np.asarray(pi, dtype=np.float32)
    sg = np.asarray(sigma, dtype=np.float32)
    if lib.coati_host_marginal_table_gtr(br_len, omega, p.ctypes.data_as(_fp), sg.ctypes.data_as(_fp),
                                         out.ctypes.data_as(_fp)) != 0:
        raise ValueError("set_subst failed")
    return out


def host_encode(anc: str, des: str):
    lib = load_library()
    ab, db = anc.encode("latin-1"), des.encode("latin-1")
    a = np.zeros(len(ab), dtype=np.uint8)
    b = np.zeros(len(db), dtype=np.uint8)
    rc = lib.coati_host_encode(ab, len(ab), db, len(db), a.ctypes.data_as(_u8p), b.ctypes.data_as(_u8p))
    if rc != 0:
        raise CoatiGpuError(rc, lib.coati_gpu_strerror(rc).decode())
    return a, b
