"""ctypes binding of libKernel.so (include/mh_kernel.h).

This is the binding a maintainer of the reference's C# wrapper would write in P/Invoke,
restated in Python for the tests and bench.py: numpy struct arrays in, numpy struct arrays
out, every call going through the exported C ABI.  There is no fallback: if the library or a
CUDA device is missing the call raises KernelError.
"""
import ctypes as C
import os

import numpy as np

from . import layout as L

_HERE = os.path.dirname(os.path.abspath(__file__))
_P = C.c_void_p

EXPORTS = ("KernelWrapper", "KernelWrapperEx", "KernelFree", "KernelLastError", "KernelEvalCosts", "KernelCreate", "KernelRun",
           "KernelRunTraced", "KernelSynchronize", "KernelResults", "KernelDeviceResults", "KernelSetStream", "KernelBest",
           "KernelStats", "KernelDestroy", "KernelDeviceInfo", "KernelBestKey", "KernelDecodeBestKey", "KernelReset", "KernelTrim",
           "KernelTemperingState", "KernelTemperingExchange", "KernelTopK", "KernelTopKDistinct", "KernelTemperingStats", "KernelShape",
           "KernelDeviceCount", "KernelTemperingLadder", "KernelTemperingProposeLadder", "KernelTemperingSetLadder")


class KernelError(RuntimeError):
    pass


def lib_path():
    """MH_LIB lets a development probe load an experimental build; the product is libKernel.so."""
    return os.environ.get("MH_LIB") or os.path.join(_HERE, "libKernel.so")


def _ptr(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_P)


def make_options(**kw):
    """mhOptions from keyword arguments.  `device=k` names CUDA ordinal k (also 0: the explicit-device flag is
    set); `devices=[...]` (or an int N = the first N ordinals) spreads the chains over several GPUs in-process."""
    o = np.zeros(1, L.mhOptions)
    o["struct_size"] = L.mhOptions.itemsize
    o["device"] = -1
    for k, v in kw.items():
        if k == "devices":
            if v is None:
                continue
            ids = list(range(v)) if isinstance(v, int) else list(v)
            if len(ids) > L.MH_MAX_DEVICES:
                raise KernelError(f"at most {L.MH_MAX_DEVICES} devices")
            o["n_devices"] = len(ids)
            o["devices"][0, :len(ids)] = ids
        elif k == "device":
            o["device"] = v
            if v is not None and v >= 0:
                o["flags"] |= L.MH_OPT_EXPLICIT_DEVICE
        else:
            o[k] = v
    return o


class Kernel:
    """Loads libKernel.so once and exposes the C ABI."""

    _lib = None

    def __init__(self):
        if Kernel._lib is None:
            path = lib_path()
            if not os.path.exists(path):
                raise KernelError(f"{path} is missing: build it with `make -C {_HERE}/csrc` (or __graft_entry__.build())")
            lib = C.CDLL(path)
            I, F, D, LL = C.c_int, C.c_float, C.c_double, C.c_longlong
            PI_, PF = C.POINTER(C.c_int), C.POINTER(C.c_float)
            room8 = [_P] * 8                                    # rss, rsa, cfg, clearances, offlimits, vertices, surfaceRectangle, srf
            # (restype, argtypes) of EVERY export: a handle is a 64-bit pointer, and ctypes would pass a bare
            # Python int as a 32-bit C int (truncating any context allocated above 4 GiB) without this table
            sig = {
                "KernelWrapper": (_P, room8 + [_P]),
                "KernelWrapperEx": (_P, room8 + [_P, _P]),
                "KernelFree": (None, [_P]),
                "KernelLastError": (C.c_char_p, []),
                "KernelEvalCosts": (I, [_P, _P, _P, I, _P, _P, _P, _P, _P, _P]),
                "KernelCreate": (_P, room8 + [I, _P]),
                "KernelRun": (I, [_P, I]),
                "KernelRunTraced": (I, [_P, I, _P]),
                "KernelSynchronize": (I, [_P]),
                "KernelResults": (I, [_P, _P, _P]),
                "KernelDeviceResults": (I, [_P, C.POINTER(_P), C.POINTER(_P)]),
                "KernelSetStream": (I, [_P, _P]),
                "KernelBest": (I, [_P, PI_, PF]),
                "KernelStats": (I, [_P, C.POINTER(D), C.POINTER(LL)]),
                "KernelDestroy": (None, [_P]),
                "KernelDeviceInfo": (I, [PI_, PI_, PI_, PI_, C.c_char_p, I]),
                "KernelBestKey": (I, [_P, _P]),
                "KernelDecodeBestKey": (None, [LL, C.POINTER(C.c_ulonglong), PF]),
                "KernelReset": (I, [_P]),
                "KernelTrim": (I, []),
                "KernelTemperingState": (I, [_P, C.POINTER(_P), C.POINTER(_P)]),
                "KernelTemperingExchange": (I, [_P, _P, _P]),
                "KernelTopK": (I, [_P, I, _P, _P]),
                "KernelTopKDistinct": (I, [_P, I, F, F, _P, _P]),
                "KernelTemperingStats": (I, [_P, _P, _P]),
                "KernelShape": (I, [_P, PI_, PI_, PI_, _P, _P]),
                "KernelDeviceCount": (I, []),
                "KernelTemperingLadder": (I, [_P, _P]),
                "KernelTemperingProposeLadder": (I, [I, _P, _P, _P, D, _P]),
                "KernelTemperingSetLadder": (I, [_P, _P]),
            }
            assert set(sig) == set(EXPORTS)
            for name, (res, args) in sig.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            Kernel._lib = lib
        self.lib = Kernel._lib

    # ---- helpers -------------------------------------------------------------------------
    def last_error(self):
        return (self.lib.KernelLastError() or b"").decode()

    def _fail(self, what):
        raise KernelError(f"{what}: {self.last_error()}")

    @staticmethod
    def _room_args(room):
        for a, dt in ((room.rss, L.relationshipStruct), (room.rsa, L.relationshipAngleStruct), (room.cfg, L.positionAndRotation),
                      (room.clearances, L.rectangle), (room.offlimits, L.rectangle), (room.vertices, L.vertex),
                      (room.surfaceRectangle, L.vertex), (room.srf, L.Surface)):
            if a.dtype != dt:                                  # e.g. np.concatenate silently repacks padded structs
                raise KernelError(f"array has dtype of itemsize {a.dtype.itemsize}, the wire struct has {dt.itemsize}")
        return [_ptr(room.rss), _ptr(room.rsa), _ptr(room.cfg), _ptr(room.clearances), _ptr(room.offlimits), _ptr(room.vertices),
                _ptr(room.surfaceRectangle), _ptr(room.srf)]

    def _unpack(self, res, n_chains, n):
        r = np.ctypeslib.as_array(C.cast(res, C.POINTER(C.c_uint8)), shape=(n_chains * L.result.itemsize,)).view(L.result)
        costs = r["costs"].copy()
        base = int(r["points"][0])
        # result[i].points must point into ONE block: base + i*n*sizeof(point) (Kernel.cu:981)
        expect = base + np.arange(n_chains, dtype=np.uint64) * np.uint64(n * L.point.itemsize)
        if not np.array_equal(r["points"], expect):
            raise KernelError("result[i].points are not slices of one block")
        pts = np.ctypeslib.as_array(C.cast(base, C.POINTER(C.c_uint8)), shape=(n_chains * n * L.point.itemsize,)).view(L.point).copy()
        self.lib.KernelFree(res)
        return pts.reshape(n_chains, n), costs

    # ---- the reference's entry point -------------------------------------------------------
    def wrapper(self, room, n_chains, iterations, block=64):
        """KernelWrapper exactly as the reference's caller invokes it (Kernel.cu:1198)."""
        g = np.zeros(1, L.gpuConfig)
        g["gridxDim"], g["blockxDim"], g["iterations"] = n_chains, block, iterations
        res = self.lib.KernelWrapper(*self._room_args(room), _ptr(g))
        if not res:
            self._fail("KernelWrapper")
        return self._unpack(res, n_chains, room.n)

    def wrapper_ex_raw(self, room, n_chains, iterations, **opts):
        """KernelWrapperEx without the copies of _unpack: returns (result* address, points view,
        costs view) over the library's malloc'd blocks; the caller must KernelFree the address."""
        g = np.zeros(1, L.gpuConfig)
        g["gridxDim"], g["blockxDim"], g["iterations"] = n_chains, 64, iterations
        o = make_options(**opts)
        res = self.lib.KernelWrapperEx(*self._room_args(room), _ptr(g), _ptr(o))
        if not res:
            self._fail("KernelWrapperEx")
        r = np.ctypeslib.as_array(C.cast(res, C.POINTER(C.c_uint8)), shape=(n_chains * L.result.itemsize,)).view(L.result)
        base = int(r["points"][0])
        pts = np.ctypeslib.as_array(C.cast(base, C.POINTER(C.c_uint8)), shape=(n_chains * room.n * L.point.itemsize,)).view(L.point)
        return res, pts.reshape(n_chains, room.n), r["costs"]

    def free(self, res):
        self.lib.KernelFree(res)

    def wrapper_ex(self, room, n_chains, iterations, **opts):
        g = np.zeros(1, L.gpuConfig)
        g["gridxDim"], g["blockxDim"], g["iterations"] = n_chains, 64, iterations
        o = make_options(**opts)
        res = self.lib.KernelWrapperEx(*self._room_args(room), _ptr(g), _ptr(o))
        if not res:
            self._fail("KernelWrapperEx")
        return self._unpack(res, n_chains, room.n)

    def eval_costs(self, room, layouts):
        n = room.n
        nl = len(layouts) // n
        if layouts.dtype != L.positionAndRotation:
            raise KernelError("layouts must have the positionAndRotation wire dtype (72-byte items)")
        out = np.zeros(nl, L.resultCosts)
        rc = self.lib.KernelEvalCosts(_ptr(room.rss), _ptr(room.rsa), _ptr(layouts), C.c_int(nl), _ptr(room.clearances),
                                      _ptr(room.offlimits), _ptr(room.vertices), _ptr(room.surfaceRectangle), _ptr(room.srf), _ptr(out))
        if rc != 0:
            self._fail("KernelEvalCosts")
        return out

    def device_info(self):
        sm, khz, ma, mi = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        name = C.create_string_buffer(256)
        if self.lib.KernelDeviceInfo(C.byref(sm), C.byref(khz), C.byref(ma), C.byref(mi), name, 256) != 0:
            self._fail("KernelDeviceInfo")
        return {"sm_count": sm.value, "sm_clock_khz": khz.value, "cc": (ma.value, mi.value), "name": name.value.decode()}

    def decode_best_key(self, key):
        g, t = C.c_ulonglong(), C.c_float()
        self.lib.KernelDecodeBestKey(C.c_longlong(int(key)), C.byref(g), C.byref(t))
        return g.value, t.value

    def device_count(self):
        n = self.lib.KernelDeviceCount()
        if n < 0:
            self._fail("KernelDeviceCount")
        return n

    def propose_ladder(self, current, attempts, accepted, damping=1.0):
        """KernelTemperingProposeLadder: the ladder that would equalise the exchange rates of neighbouring rungs."""
        cur = np.ascontiguousarray(current, np.float64)
        att = np.ascontiguousarray(attempts, np.int64)
        acc = np.ascontiguousarray(accepted, np.int64)
        out = np.zeros(len(cur), np.float64)
        if self.lib.KernelTemperingProposeLadder(len(cur), _ptr(cur), _ptr(att), _ptr(acc), float(damping), _ptr(out)) != 0:
            self._fail("KernelTemperingProposeLadder")
        return out

    def create(self, room, n_chains, **opts):
        return Context(self, room, n_chains, **opts)


class Context:
    """Persistent device-resident run (KernelCreate .. KernelDestroy)."""

    def __init__(self, k, room, n_chains, **opts):
        self.k, self.room, self.n_chains = k, room, n_chains
        o = make_options(**opts)
        self.h = k.lib.KernelCreate(*Kernel._room_args(room), C.c_int(n_chains), _ptr(o))
        if not self.h:
            k._fail("KernelCreate")

    def run(self, iterations):
        if self.k.lib.KernelRun(self.h, iterations) != 0:
            self.k._fail("KernelRun")

    def run_traced(self, iterations):
        tr = np.zeros(iterations * self.n_chains, L.mhTraceEntry)
        if self.k.lib.KernelRunTraced(self.h, iterations, _ptr(tr)) != 0:
            self.k._fail("KernelRunTraced")
        return tr.reshape(iterations, self.n_chains)

    def synchronize(self):
        if self.k.lib.KernelSynchronize(self.h) != 0:
            self.k._fail("KernelSynchronize")

    def results(self, points=None, costs=None):
        n = self.room.n
        pts = np.zeros(self.n_chains * n, L.point) if points is None else points
        cs = np.zeros(self.n_chains, L.resultCosts) if costs is None else costs
        if self.k.lib.KernelResults(self.h, _ptr(pts), _ptr(cs)) != 0:
            self.k._fail("KernelResults")
        return pts.reshape(self.n_chains, n), cs

    def device_results(self):
        dp, dc = _P(), _P()
        if self.k.lib.KernelDeviceResults(self.h, C.byref(dp), C.byref(dc)) != 0:
            self.k._fail("KernelDeviceResults")
        return dp.value, dc.value

    def set_stream(self, stream_handle):
        """stream_handle: a cudaStream_t as an integer (e.g. torch.cuda.current_stream().cuda_stream).
        torch reports the legacy default stream as 0; the C ABI names it cudaStreamLegacy = 0x1."""
        if not stream_handle:
            stream_handle = 1
        if self.k.lib.KernelSetStream(self.h, _P(stream_handle)) != 0:
            self.k._fail("KernelSetStream")
        self._stream_handle = stream_handle if stream_handle != 1 else 0

    def best(self):
        i, t = C.c_int(), C.c_float()
        if self.k.lib.KernelBest(self.h, C.byref(i), C.byref(t)) != 0:
            self.k._fail("KernelBest")
        return i.value, t.value

    def best_key(self, d_key):
        """d_key: device address of one int64 (e.g. a torch tensor's data_ptr())."""
        if self.k.lib.KernelBestKey(self.h, _P(d_key)) != 0:
            self.k._fail("KernelBestKey")

    def tempering_state(self):
        """Device addresses of this context's per-chain totalCosts and betas (float32[n_chains])."""
        dt, db = _P(), _P()
        if self.k.lib.KernelTemperingState(self.h, C.byref(dt), C.byref(db)) != 0:
            self.k._fail("KernelTemperingState")
        return dt.value, db.value

    def tempering_exchange(self, d_all_totals, d_all_betas):
        if self.k.lib.KernelTemperingExchange(self.h, _P(d_all_totals), _P(d_all_betas)) != 0:
            self.k._fail("KernelTemperingExchange")

    def tempering_stats(self, rungs):
        """(attempts, accepted) per pair of neighbouring rungs, as counted by this context's chains."""
        att = np.zeros(rungs - 1, np.int64)
        acc = np.zeros(rungs - 1, np.int64)
        if self.k.lib.KernelTemperingStats(self.h, _ptr(att), _ptr(acc)) != rungs - 1:
            self.k._fail("KernelTemperingStats")
        return att, acc

    def ladder(self, rungs):
        out = np.zeros(rungs, np.float64)
        if self.k.lib.KernelTemperingLadder(self.h, _ptr(out)) != rungs:
            self.k._fail("KernelTemperingLadder")
        return out

    def set_ladder(self, betas):
        b = np.ascontiguousarray(betas, np.float64)
        if self.k.lib.KernelTemperingSetLadder(self.h, _ptr(b)) != 0:
            self.k._fail("KernelTemperingSetLadder")

    def reset(self):
        if self.k.lib.KernelReset(self.h) != 0:
            self.k._fail("KernelReset")

    def top_k(self, k):
        idx = np.zeros(k, np.int32)
        tot = np.zeros(k, np.float32)
        m = self.k.lib.KernelTopK(self.h, C.c_int(k), _ptr(idx), _ptr(tot))
        if m < 0:
            self.k._fail("KernelTopK")
        return idx[:m], tot[:m]

    def top_k_distinct(self, k, min_distance, rot_weight=0.0):
        idx = np.zeros(k, np.int32)
        tot = np.zeros(k, np.float32)
        m = self.k.lib.KernelTopKDistinct(self.h, C.c_int(k), C.c_float(min_distance), C.c_float(rot_weight), _ptr(idx), _ptr(tot))
        if m < 0:
            self.k._fail("KernelTopKDistinct")
        return idx[:m], tot[:m]

    def shape(self):
        """How the context was laid out: lanes per chain, evaluation form, and (device, chains) per device."""
        lanes, form, nd = C.c_int(), C.c_int(), C.c_int()
        dev = np.zeros(L.MH_MAX_DEVICES, np.int32)
        cnt = np.zeros(L.MH_MAX_DEVICES, np.int32)
        if self.k.lib.KernelShape(self.h, C.byref(lanes), C.byref(form), C.byref(nd), _ptr(dev), _ptr(cnt)) != 0:
            self.k._fail("KernelShape")
        return {"lanes_per_chain": lanes.value, "eval_form": form.value, "devices": [int(d) for d in dev[:nd.value]],
                "chains_per_device": [int(c) for c in cnt[:nd.value]]}

    def stats(self):
        ms, n = C.c_double(), C.c_longlong()
        if self.k.lib.KernelStats(self.h, C.byref(ms), C.byref(n)) != 0:
            self.k._fail("KernelStats")
        return ms.value, n.value

    def close(self):
        if self.h:
            self.k.lib.KernelDestroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
