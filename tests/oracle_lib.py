"""ctypes access to the test oracle (oracle/liboracle.so) and, when present, to the
reference-built libraries under oracle/_ref/.  Test infrastructure: imported by tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs only."""
import ctypes as C
import importlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mh = importlib.import_module("metropolis-hastings-gpgpu_b200.layout")

oracleRunOptions = np.dtype(
    {"names": ["seed", "chain_offset", "iteration_offset", "beta_start", "beta_end", "schedule", "schedule_length", "result_mode",
               "with_offlimits", "threads", "tempering_rungs", "exchange_interval", "_pad"],
     "formats": ["<u8", "<u8", "<u8", "<f8", "<f8", "<i4", "<i4", "<i4", "<i4", "<i4", "<i4", "<i4", "<i4"],
     "offsets": [0, 8, 16, 24, 32, 40, 44, 48, 52, 56, 60, 64, 68], "itemsize": 72})

_P = C.c_void_p


def _ptr(a):
    return a.ctypes.data_as(_P) if a is not None else None


def build_oracle():
    import subprocess
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)


def build_native_oracle():
    """The same mh_oracle.c compiled for THIS host's CPU (-march=native), for the timed CPU baseline
    (BASELINE.md section 2(b)); the portable build stays the checker.  Returns a path or None."""
    import subprocess
    import tempfile
    out = os.path.join(tempfile.gettempdir(), f"liboracle_native_{os.getuid()}.so")
    src = os.path.join(ROOT, "oracle", "mh_oracle.c")
    try:
        subprocess.run(["gcc", "-std=c11", "-O2", "-march=native", "-ffp-contract=off", "-fopenmp", "-fvisibility=hidden", "-shared",
                        "-fPIC", "-o", out, src, "-lm"], check=True, capture_output=True, timeout=120)
        return out
    except Exception:
        return None


class Oracle:
    def __init__(self, path=None):
        path = path or os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        self.lib = C.CDLL(path)
        self.lib.oracle_run_timed.restype = C.c_double
        self.lib.oracle_angle_branch_margin.restype = C.c_double
        self.lib.oracle_uniform.restype = C.c_float
        self.lib.oracle_uniform.argtypes = [C.c_uint32]
        self.lib.oracle_random_int.argtypes = [C.c_float, C.c_int, C.c_int]

    def philox(self, ctr, key):
        c = np.asarray(ctr, np.uint32)
        k = np.asarray(key, np.uint32)
        o = np.zeros(4, np.uint32)
        self.lib.oracle_philox4x32_10(_ptr(c), _ptr(k), _ptr(o))
        return o

    def costs(self, room, cfg=None, raw=False):
        cfg = room.cfg if cfg is None else cfg
        assert cfg.dtype == mh.positionAndRotation, "layout array lost the 72-byte wire dtype"
        out = np.zeros(1, mh.resultCosts)
        raw8 = np.zeros(8, np.float64)
        self.lib.oracle_costs(_ptr(room.srf), _ptr(cfg), _ptr(room.rss), _ptr(room.rsa), _ptr(room.vertices), _ptr(room.clearances),
                              _ptr(room.offlimits), _ptr(room.surfaceRectangle), _ptr(out), _ptr(raw8))
        return (out[0], raw8) if raw else out[0]

    def costs_batch(self, room, layouts):
        n = room.n
        nl = len(layouts) // n
        assert layouts.dtype == mh.positionAndRotation, "layout array lost the 72-byte wire dtype"
        out = np.zeros(nl, mh.resultCosts)
        self.lib.oracle_costs_batch(_ptr(room.srf), _ptr(layouts), C.c_int(nl), _ptr(room.rss), _ptr(room.rsa), _ptr(room.vertices),
                                    _ptr(room.clearances), _ptr(room.offlimits), _ptr(room.surfaceRectangle), _ptr(out))
        return out

    def angle_margin(self, room, cfg):
        return float(self.lib.oracle_angle_branch_margin(_ptr(room.srf), _ptr(cfg), _ptr(room.rsa)))

    def run(self, room, n_chains, iterations, seed=1, chain_offset=0, iteration_offset=0, beta_start=0.0, beta_end=0.0, schedule=0,
            schedule_length=0, result_mode=0, with_offlimits=1, threads=0, tempering_rungs=0, exchange_interval=0, trace=False,
            timed=False):
        o = np.zeros(1, oracleRunOptions)
        o["seed"], o["chain_offset"], o["iteration_offset"] = seed, chain_offset, iteration_offset
        o["beta_start"], o["beta_end"], o["schedule"], o["schedule_length"] = beta_start, beta_end, schedule, schedule_length
        o["result_mode"], o["with_offlimits"], o["threads"] = result_mode, with_offlimits, threads
        o["tempering_rungs"], o["exchange_interval"] = tempering_rungs, exchange_interval
        pts = np.zeros(n_chains * room.n, mh.point)
        costs = np.zeros(n_chains, mh.resultCosts)
        tr = np.zeros(iterations * n_chains, mh.mhTraceEntry) if trace else None
        args = [_ptr(room.srf), _ptr(room.rss), _ptr(room.rsa), _ptr(room.cfg), _ptr(room.clearances), _ptr(room.offlimits),
                _ptr(room.vertices), _ptr(room.surfaceRectangle), C.c_int(n_chains), C.c_int(iterations), _ptr(o), _ptr(pts), _ptr(costs)]
        if timed:
            th = C.c_int(0)
            secs = self.lib.oracle_run_timed(*args, C.byref(th))
            return pts.reshape(n_chains, room.n), costs, float(secs), th.value
        rc = self.lib.oracle_run(*args, _ptr(tr))
        if rc < 0:
            raise RuntimeError("oracle_run rejected the arguments")
        if trace:
            return pts.reshape(n_chains, room.n), costs, tr.reshape(iterations, n_chains)
        return pts.reshape(n_chains, room.n), costs


def ref_host_path():
    return os.path.join(ROOT, "oracle", "_ref", "libref_costs_host.so")


def ref_gpu_path():
    return os.path.join(ROOT, "oracle", "_ref", "libKernel_ref.so")


class RefHost:
    """The reference's own Costs() compiled as host C++ (oracle/_ref/libref_costs_host.so)."""

    def __init__(self):
        self.lib = C.CDLL(ref_host_path())

    def costs(self, room, cfg=None, raw=False):
        cfg = room.cfg if cfg is None else cfg
        out = np.zeros(1, mh.resultCosts)
        raw8 = np.zeros(8, np.float64)
        self.lib.ref_costs(_ptr(room.srf), _ptr(cfg), _ptr(room.rss), _ptr(room.rsa), _ptr(room.vertices), _ptr(room.clearances),
                           _ptr(room.offlimits), _ptr(room.surfaceRectangle), _ptr(out), _ptr(raw8))
        return (out[0], raw8) if raw else out[0]

    def sizes(self):
        s = np.zeros(10, np.int32)
        self.lib.ref_sizes(_ptr(s))
        return s


class RefGPU:
    """The reference kernel rebuilt for sm_100 (oracle/_ref/libKernel_ref.so); needs a GPU."""

    def __init__(self):
        self.lib = C.CDLL(ref_gpu_path())
        self.lib.RefTimedWrapper.restype = C.c_void_p
        self.lib.RefSetHeap.argtypes = [C.c_size_t]
        self.lib.RefProposeGPU.argtypes = [_P, _P, C.c_int, _P, C.c_uint]
        self.lib.RefAcceptGPU.argtypes = [_P, _P, C.c_int, C.c_uint, _P]
        self.lib.RefInitRngMs.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float)]

    def set_heap(self, nbytes):
        return self.lib.RefSetHeap(nbytes)

    def costs_gpu(self, room, layouts):
        nl = len(layouts) // room.n
        out = np.zeros(nl, mh.resultCosts)
        rc = self.lib.RefCostsGPU(_ptr(room.srf), _ptr(layouts), C.c_int(nl), _ptr(room.rss), _ptr(room.rsa), _ptr(room.vertices),
                                  _ptr(room.clearances), _ptr(room.offlimits), _ptr(room.surfaceRectangle), _ptr(out))
        if rc != 0:
            raise RuntimeError(f"RefCostsGPU: cuda error {rc}")
        return out

    def propose_gpu(self, room, layouts, seed):
        """The reference's own propose() (Kernel.cu:576-704) applied once to every layout, one thread and
        one XORWOW state (seeded like initRNG, Kernel.cu:159) per layout.  Returns the mutated copies."""
        out = np.ascontiguousarray(layouts).copy()
        assert out.dtype == mh.positionAndRotation
        rc = self.lib.RefProposeGPU(_ptr(room.srf), _ptr(out), C.c_int(len(out) // room.n), _ptr(room.surfaceRectangle), C.c_uint(seed))
        if rc != 0:
            raise RuntimeError(f"RefProposeGPU: cuda error {rc}")
        return out

    def accept_gpu(self, star, cur, seed):
        """The reference's own Accept() (Kernel.cu:706-713) on arrays of (costStar, costCur)."""
        a = np.ascontiguousarray(star, np.float64)
        b = np.ascontiguousarray(cur, np.float64)
        out = np.zeros(len(a), np.int32)
        rc = self.lib.RefAcceptGPU(_ptr(a), _ptr(b), C.c_int(len(a)), C.c_uint(seed), _ptr(out))
        if rc != 0:
            raise RuntimeError(f"RefAcceptGPU: cuda error {rc}")
        return out

    def init_rng_ms(self, grid, block):
        """Device milliseconds of the reference's initRNG launch alone (Kernel.cu:939-943)."""
        ms = C.c_float(0)
        rc = self.lib.RefInitRngMs(C.c_int(grid), C.c_int(block), C.byref(ms))
        if rc != 0:
            raise RuntimeError(f"RefInitRngMs: cuda error {rc}")
        return float(ms.value)

    def run(self, room, n_chains, iterations, block=64):
        """The reference's KernelWrapper, unmodified.  Returns (points[n_chains, n], device ms of
        the whole call).  Its `costs` are uninitialised memory (quirk Q3) and are not returned."""
        g = np.zeros(1, mh.gpuConfig)
        g["gridxDim"], g["blockxDim"], g["iterations"] = n_chains, block, iterations
        ms = C.c_float(0)
        res = self.lib.RefTimedWrapper(_ptr(room.rss), _ptr(room.rsa), _ptr(room.cfg), _ptr(room.clearances), _ptr(room.offlimits),
                                       _ptr(room.vertices), _ptr(room.surfaceRectangle), _ptr(room.srf), _ptr(g), C.byref(ms))
        if not res:
            raise RuntimeError("reference KernelWrapper returned NULL")
        r = np.ctypeslib.as_array(C.cast(res, C.POINTER(C.c_uint8)), shape=(n_chains * 40,)).view(mh.result)
        base = int(r["points"][0])
        pts = np.ctypeslib.as_array(C.cast(base, C.POINTER(C.c_uint8)), shape=(n_chains * room.n * 24,)).view(mh.point).copy()
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        libc.free(base)
        libc.free(res)
        return pts.reshape(n_chains, room.n), float(ms.value)
