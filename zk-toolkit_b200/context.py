"""Device context and device-resident point sets over the C ABI."""
import ctypes

import numpy as np

from . import _lib as L


class PointSet:
    """A `&[G1Point]` / `&[G2Point]` slice kept resident on one B200 (e.g. a CRS vector)."""

    def __init__(self, ctx, handle, group, n, precomputed):
        self.ctx, self.handle, self.group, self.n, self.precomputed = ctx, handle, group, n, precomputed

    def __len__(self):
        return self.n

    def info(self):
        c, w, pre, sub = ctypes.c_uint(0), ctypes.c_uint(0), ctypes.c_int(0), ctypes.c_int(0)
        self.ctx._check(self.ctx.lib.zkmsm_points_info(self.handle, ctypes.byref(c), ctypes.byref(w), ctypes.byref(pre),
                                                       ctypes.byref(sub)))
        return {"c": c.value, "windows": w.value, "precomputed": bool(pre.value), "subgroup": bool(sub.value)}

    def read(self, first=0, n=None):
        """canonical affine limbs and infinity flags of points [first, first+n)"""
        n = self.n - first if n is None else n
        words = L.G1_WORDS if self.group == 1 else L.G2_WORDS
        xy = np.zeros((n, words), dtype=np.uint32)
        inf = np.zeros(n, dtype=np.uint8)
        self.ctx._check(self.ctx.lib.zkmsm_points_read(self.ctx.h, self.handle, first, n, L.dptr(xy), L.dptr(inf)))
        return xy, inf

    def free(self):
        if self.handle is not None:
            self.ctx.lib.zkmsm_points_free(self.ctx.h, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One CUDA device.  Raises ZkmsmError when no sm_100 GPU is usable: there is no CPU path."""

    def __init__(self, device=0):
        self.lib = L.load()
        h = ctypes.c_void_p()
        rc = self.lib.zkmsm_create(device, ctypes.byref(h))
        if rc != L.OK:
            raise L.ZkmsmError(rc, "zkmsm_create failed (an sm_100 CUDA device is required)")
        self.h = h
        self.device = device

    def close(self):
        for p in getattr(self, "_pinned", []):
            self.lib.zkmsm_host_free(p)
        self._pinned = []
        if getattr(self, "h", None):
            self.lib.zkmsm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != L.OK:
            raise L.ZkmsmError(rc, self.lib.zkmsm_last_error(self.h).decode())

    def pinned_array(self, shape, dtype=np.uint32):
        """numpy array in page-locked host memory (zkmsm_host_alloc): H2D copies of it run at full PCIe rate
        and asynchronously.  Freed with the context."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = ctypes.c_void_p()
        rc = self.lib.zkmsm_host_alloc(max(nbytes, 16), ctypes.byref(p))
        if rc != L.OK:
            raise L.ZkmsmError(rc, "zkmsm_host_alloc")
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)
        buf = (ctypes.c_uint8 * max(nbytes, 16)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.zkmsm_set_stream(self.h, ctypes.c_void_p(cuda_stream_ptr or 0)))

    def set_window(self, c):
        self._check(self.lib.zkmsm_set_window(self.h, c))

    def set_option(self, name, value):
        """tuning / cross-check switch of this context (zkmsm_set_option; the ZKMSM_* environment is read once, at creation)"""
        self._check(self.lib.zkmsm_set_option(self.h, name.encode(), int(value)))

    # ---- point sets
    def _words(self, group):
        return L.G1_WORDS if group == 1 else L.G2_WORDS

    @staticmethod
    def _flags(precompute, in_subgroup, check_subgroup=False):
        return (L.PRECOMPUTE if precompute else 0) | (L.SUBGROUP if in_subgroup else 0) | (L.CHECK_SUBGROUP if check_subgroup else 0)

    def load_points(self, group, xy, inf=None, precompute=False, in_subgroup=False, check_subgroup=False):
        xy = L.as_u32(xy, self._words(group)).reshape(-1, self._words(group))
        n = xy.shape[0]
        if inf is not None:
            inf = np.ascontiguousarray(inf, dtype=np.uint8)
            assert inf.shape == (n,)
        fn = self.lib.zkmsm_g1_load_points if group == 1 else self.lib.zkmsm_g2_load_points
        h = ctypes.c_void_p()
        self._check(fn(self.h, L.dptr(xy), L.dptr(inf), n, self._flags(precompute, in_subgroup, check_subgroup), ctypes.byref(h)))
        return PointSet(self, h, group, n, precompute)

    def points_from_scalars(self, group, base_xy, scalars, precompute=False, in_subgroup=False):
        base = L.as_u32(base_xy, self._words(group)).reshape(-1)
        sc = L.as_u32(scalars, 8).reshape(-1, 8)
        fn = self.lib.zkmsm_g1_points_from_scalars if group == 1 else self.lib.zkmsm_g2_points_from_scalars
        h = ctypes.c_void_p()
        self._check(fn(self.h, L.dptr(base), L.dptr(sc), sc.shape[0], self._flags(precompute, in_subgroup), ctypes.byref(h)))
        return PointSet(self, h, group, sc.shape[0], precompute)

    def mul_base(self, group, base_xy, scalars):
        base = L.as_u32(base_xy, self._words(group)).reshape(-1)
        sc = L.as_u32(scalars, 8).reshape(-1, 8)
        n = sc.shape[0]
        out = np.zeros((n, self._words(group)), dtype=np.uint32)
        inf = np.zeros(n, dtype=np.uint8)
        fn = self.lib.zkmsm_g1_mul_base if group == 1 else self.lib.zkmsm_g2_mul_base
        self._check(fn(self.h, L.dptr(base), L.dptr(sc), n, L.dptr(out), L.dptr(inf)))
        return out, inf

    # ---- MSM
    def msm(self, pts: PointSet, scalars, n=None):
        """sum_{i<n} scalars[i] * pts[i]; scalars: host (n, 8) uint32.  Returns (xy limbs, is_inf)."""
        sc = L.as_u32(scalars, 8).reshape(-1, 8) if len(scalars) else np.zeros((0, 8), dtype=np.uint32)
        n = sc.shape[0] if n is None else n
        out = np.zeros(self._words(pts.group), dtype=np.uint32)
        inf = ctypes.c_int(0)
        fn = self.lib.zkmsm_g1_msm if pts.group == 1 else self.lib.zkmsm_g2_msm
        self._check(fn(self.h, pts.handle, L.dptr(sc), n, L.dptr(out), ctypes.byref(inf)))
        return out, bool(inf.value)

    def msm_host_ptr(self, pts: PointSet, scalars_ptr, n):
        """same, scalars given as a raw host pointer (e.g. pinned memory)"""
        out = np.zeros(self._words(pts.group), dtype=np.uint32)
        inf = ctypes.c_int(0)
        fn = self.lib.zkmsm_g1_msm if pts.group == 1 else self.lib.zkmsm_g2_msm
        self._check(fn(self.h, pts.handle, ctypes.c_void_p(scalars_ptr), n, L.dptr(out), ctypes.byref(inf)))
        return out, bool(inf.value)

    def msm_device(self, pts: PointSet, scalars_dev_ptr, n):
        out = np.zeros(self._words(pts.group), dtype=np.uint32)
        inf = ctypes.c_int(0)
        fn = self.lib.zkmsm_g1_msm_device if pts.group == 1 else self.lib.zkmsm_g2_msm_device
        self._check(fn(self.h, pts.handle, ctypes.c_void_p(scalars_dev_ptr), n, L.dptr(out), ctypes.byref(inf)))
        return out, bool(inf.value)

    def msm_begin(self, pts: PointSet, scalars, n=None):
        """start an MSM with host scalars (keep the array alive until msm_result); returns immediately"""
        sc = L.as_u32(scalars, 8).reshape(-1, 8)
        n = sc.shape[0] if n is None else n
        fn = self.lib.zkmsm_g1_msm_begin if pts.group == 1 else self.lib.zkmsm_g2_msm_begin
        self._check(fn(self.h, pts.handle, L.dptr(sc), n))
        self._inflight = sc

    def msm_begin_ptr(self, pts: PointSet, scalars_ptr, n):
        """msm_begin with a raw host pointer (e.g. pinned memory, which makes the copy asynchronous)"""
        fn = self.lib.zkmsm_g1_msm_begin if pts.group == 1 else self.lib.zkmsm_g2_msm_begin
        self._check(fn(self.h, pts.handle, ctypes.c_void_p(scalars_ptr), n))

    def msm_enqueue(self, pts: PointSet, scalars_dev_ptr, n):
        fn = self.lib.zkmsm_g1_msm_enqueue if pts.group == 1 else self.lib.zkmsm_g2_msm_enqueue
        self._check(fn(self.h, pts.handle, ctypes.c_void_p(scalars_dev_ptr), n))

    def msm_result(self, group):
        out = np.zeros(self._words(group), dtype=np.uint32)
        inf = ctypes.c_int(0)
        fn = self.lib.zkmsm_g1_msm_result if group == 1 else self.lib.zkmsm_g2_msm_result
        self._check(fn(self.h, L.dptr(out), ctypes.byref(inf)))
        return out, bool(inf.value)

    def last_launch_count(self):
        return self.lib.zkmsm_last_launch_count(self.h)

    def profile(self, enable=True):
        self._check(self.lib.zkmsm_profile(self.h, 1 if enable else 0))

    def profile_read(self):
        """[(kernel name, ms, logical threads)] of the last MSM, in launch order"""
        cap, stride = 256, 32
        names = ctypes.create_string_buffer(cap * stride)
        ms = np.zeros(cap, dtype=np.float32)
        thr = np.zeros(cap, dtype=np.uint32)
        n = self.lib.zkmsm_profile_read(self.h, cap, names, stride, L.dptr(ms), L.dptr(thr))
        if n < 0:
            self._check(n)
        raw = names.raw
        return [(raw[i * stride:(i + 1) * stride].split(b"\0")[0].decode(), float(ms[i]), int(thr[i])) for i in range(n)]

    def msm_oneshot(self, group, xy, inf, scalars):
        xy = L.as_u32(xy, self._words(group)).reshape(-1, self._words(group))
        sc = L.as_u32(scalars, 8).reshape(-1, 8) if len(scalars) else np.zeros((0, 8), dtype=np.uint32)
        if inf is not None:
            inf = np.ascontiguousarray(inf, dtype=np.uint8)
        if sc.shape[0] > xy.shape[0]:
            raise L.ZkmsmError(-5, "more scalars than points")
        out = np.zeros(self._words(group), dtype=np.uint32)
        oinf = ctypes.c_int(0)
        fn = self.lib.zkmsm_g1_msm_oneshot if group == 1 else self.lib.zkmsm_g2_msm_oneshot
        self._check(fn(self.h, L.dptr(xy), L.dptr(inf), L.dptr(sc), sc.shape[0], L.dptr(out), ctypes.byref(oinf)))
        return out, bool(oinf.value)

    # ---- multi-GPU partials
    def msm_partial(self, pts: PointSet, scalars, n=None):
        sc = L.as_u32(scalars, 8).reshape(-1, 8) if len(scalars) else np.zeros((0, 8), dtype=np.uint32)
        n = sc.shape[0] if n is None else n
        words = L.G1_PARTIAL_WORDS if pts.group == 1 else L.G2_PARTIAL_WORDS
        out = np.zeros(words, dtype=np.uint32)
        fn = self.lib.zkmsm_g1_msm_partial if pts.group == 1 else self.lib.zkmsm_g2_msm_partial
        self._check(fn(self.h, pts.handle, L.dptr(sc), n, L.dptr(out)))
        return out

    def msm_partial_device(self, pts: PointSet, scalars_dev_ptr, n, out_dev_ptr):
        fn = self.lib.zkmsm_g1_msm_partial_device if pts.group == 1 else self.lib.zkmsm_g2_msm_partial_device
        self._check(fn(self.h, pts.handle, ctypes.c_void_p(scalars_dev_ptr), n, ctypes.c_void_p(out_dev_ptr)))

    def msm_partial_range(self, pts: PointSet, scalars, rank, world, n=None):
        """this rank's share of the MSM under the bucket-range split (whole precomputed set on every device)"""
        sc = L.as_u32(scalars, 8).reshape(-1, 8) if len(scalars) else np.zeros((0, 8), dtype=np.uint32)
        n = sc.shape[0] if n is None else n
        words = L.G1_PARTIAL_WORDS if pts.group == 1 else L.G2_PARTIAL_WORDS
        out = np.zeros(words, dtype=np.uint32)
        fn = self.lib.zkmsm_g1_msm_partial_range if pts.group == 1 else self.lib.zkmsm_g2_msm_partial_range
        self._check(fn(self.h, pts.handle, L.dptr(sc), n, rank, world, L.dptr(out)))
        return out

    def msm_partial_range_device(self, pts: PointSet, scalars_dev_ptr, n, rank, world, out_dev_ptr):
        fn = self.lib.zkmsm_g1_msm_partial_range_device if pts.group == 1 else self.lib.zkmsm_g2_msm_partial_range_device
        self._check(fn(self.h, pts.handle, ctypes.c_void_p(scalars_dev_ptr), n, rank, world, ctypes.c_void_p(out_dev_ptr)))

    def combine(self, group, partials):
        words = L.G1_PARTIAL_WORDS if group == 1 else L.G2_PARTIAL_WORDS
        parts = L.as_u32(partials, words).reshape(-1, words)
        out = np.zeros(self._words(group), dtype=np.uint32)
        inf = ctypes.c_int(0)
        fn = self.lib.zkmsm_g1_combine if group == 1 else self.lib.zkmsm_g2_combine
        self._check(fn(self.h, L.dptr(parts), parts.shape[0], L.dptr(out), ctypes.byref(inf)))
        return out, bool(inf.value)

    def combine_device(self, partials_dev_ptr, k, group=1):
        out = np.zeros(self._words(group), dtype=np.uint32)
        inf = ctypes.c_int(0)
        fn = self.lib.zkmsm_g1_combine_device if group == 1 else self.lib.zkmsm_g2_combine_device
        self._check(fn(self.h, ctypes.c_void_p(partials_dev_ptr), k, L.dptr(out), ctypes.byref(inf)))
        return out, bool(inf.value)

    def combine_enqueue(self, partials_dev_ptr, k, group=1):
        """stream-ordered combine of k device-resident partials; fetch with msm_result(group)"""
        fn = self.lib.zkmsm_g1_combine_enqueue if group == 1 else self.lib.zkmsm_g2_combine_enqueue
        self._check(fn(self.h, ctypes.c_void_p(partials_dev_ptr), k))

    def fr_aggregate(self, polys, wires):
        """out[j] = sum_i wires[i] * polys[i][j] mod r; polys (n_wires, n, 8) uint32, wires (n_wires, 8) -> (n, 8)"""
        polys = L.as_u32(polys, 8)
        wires = L.as_u32(wires, 8).reshape(-1, 8)
        n_wires, n = polys.shape[0], polys.shape[1]
        assert polys.ndim == 3 and wires.shape[0] == n_wires
        out = np.zeros((n, 8), dtype=np.uint32)
        self._check(self.lib.zkmsm_fr_aggregate(self.h, L.dptr(polys), n_wires, n, L.dptr(wires), L.dptr(out)))
        return out

    def fr_quotient(self, u, v, w):
        """h = (u v - w) / prod_{k=1..n} (x - k); u, v, w: (n, 8) uint32 coefficient arrays.  Returns ((n-1, 8) array,
        exact flag)."""
        u, v, w = (L.as_u32(a, 8).reshape(-1, 8) for a in (u, v, w))
        n = u.shape[0]
        assert v.shape[0] == n and w.shape[0] == n
        out = np.zeros((n - 1, 8), dtype=np.uint32)
        exact = ctypes.c_int(0)
        self._check(self.lib.zkmsm_fr_quotient(self.h, L.dptr(u), L.dptr(v), L.dptr(w), n, L.dptr(out), ctypes.byref(exact)))
        return out, bool(exact.value)

    def bench_imad(self, variant, iters=4096):
        lp, ms = ctypes.c_double(0), ctypes.c_double(0)
        self._check(self.lib.zkmsm_bench_imad(self.h, variant, iters, ctypes.byref(lp), ctypes.byref(ms)))
        return lp.value, ms.value
