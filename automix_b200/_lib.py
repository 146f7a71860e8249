"""ctypes binding of the C-ABI in include/amx.h (automix_b200/lib/libautomix.so).

This is the host-side harness used by tests/ and bench.py; the product boundary is the C-ABI
itself (and the drop-in LibAutoMix C API built on it).  There is no fallback: if the shared
library is missing or no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AMX_LIB_PATH") or os.path.join(HERE, "lib", "libautomix.so")  # (override: build experiments)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_u64p = C.POINTER(C.c_ulonglong)


class AmxError(RuntimeError):
    pass


class RjStats(C.Structure):
    _fields_ = [("acc_block", C.c_ulonglong), ("try_block", C.c_ulonglong),
                ("acc_single", C.c_ulonglong), ("try_single", C.c_ulonglong),
                ("acc_jump", C.c_ulonglong), ("try_jump", C.c_ulonglong),
                ("flops", C.c_ulonglong), ("draws", C.c_ulonglong), ("kernel_ms", C.c_double)]


class EmResult(C.Structure):
    _fields_ = [("L", C.c_int), ("iters", C.c_int), ("status", C.c_int), ("comp_steps", C.c_long),
                ("kernel_ms", C.c_double), ("flops", C.c_double), ("bytes", C.c_double), ("bytes_requested", C.c_double)]


_lib = None

# every symbol include/amx.h declares (tests check that the library exports them all)
EXPORTS = [
    "amx_last_error", "amx_version", "amx_device_count", "amx_set_device", "amx_set_stream",
    "amx_synchronize", "amx_set_deferred_sync", "amx_release_workspace", "amx_launch_count", "amx_measure_fp64_peak", "amx_mix_logpdf",
    "amx_mix_logpdf_dev", "amx_target_gaussmix", "amx_target_quad", "amx_target_coalmine", "amx_target_mixnorm", "amx_target_plugin",
    "amx_target_host_scalar", "amx_target_host_batched", "amx_target_destroy", "amx_target_eval",
    "amx_proposal_create", "amx_proposal_destroy", "amx_rj_create", "amx_rj_destroy",
    "amx_rj_set_tape", "amx_rj_set_sort", "amx_rj_set_pk_mode", "amx_rj_get_pk_shared", "amx_rj_set_chain_base", "amx_rj_set_modes", "amx_rwm_set_dof", "amx_copy_dev", "amx_rj_init_chains", "amx_rj_set_state", "amx_rj_get_state", "amx_rj_sweeps",
    "amx_rj_collect", "amx_rj_visit_se", "amx_rj_get_trace", "amx_rj_visits_dev", "amx_sokal", "amx_sokal_dev", "amx_rj_sokal",
    "amx_rj_moments_reset", "amx_rj_moments_accumulate", "amx_rj_moments_get", "amx_em_fit", "amx_em_fit_dev",
    "amx_em_draw_init", "amx_em_fit_multi", "amx_autorj_fit", "amx_rwm_adapt", "amx_rwm_adapt_all", "amx_fam_plan", "amx_fam_pack",
]


def sokal(x):
    """Sokal integrated autocorrelation time of each row of ``x`` ([nseries, n] or [n]); see amx_sokal."""
    x = np.atleast_2d(f64(x))
    ns, n = x.shape
    var, tau, m = np.zeros(ns), np.zeros(ns), np.zeros(ns, np.int32)
    check(lib().amx_sokal(ns, n, _d(x), _d(var), _d(tau), _i(m)))
    return var, tau, m


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AmxError(f"{LIB_PATH} is missing: run `python -m automix_b200.build` (no fallback exists)")
    L = C.CDLL(LIB_PATH)
    L.amx_last_error.restype = C.c_char_p
    L.amx_version.restype = C.c_char_p
    L.amx_set_stream.argtypes = [C.c_void_p]
    L.amx_launch_count.restype = C.c_ulonglong
    L.amx_launch_count.argtypes = [C.c_int]
    L.amx_measure_fp64_peak.argtypes = [_dp]
    L.amx_mix_logpdf.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_long, _dp, _dp, _dp]
    L.amx_mix_logpdf_dev.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]
    L.amx_target_gaussmix.restype = C.c_void_p
    L.amx_target_gaussmix.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, C.c_int]
    L.amx_target_quad.restype = C.c_void_p
    L.amx_target_quad.argtypes = [C.c_int, _ip, _dp, _dp, _dp, _dp]
    L.amx_target_coalmine.restype = C.c_void_p
    L.amx_target_plugin.restype = C.c_void_p
    L.amx_target_plugin.argtypes = [C.c_char_p, C.c_int, _ip, C.c_void_p, C.c_size_t, C.c_int]
    L.amx_target_mixnorm.restype = C.c_void_p
    L.amx_target_mixnorm.argtypes = [C.c_int, _ip, C.c_int, _dp, _dp]
    L.amx_target_host_scalar.restype = C.c_void_p
    L.amx_target_host_scalar.argtypes = [C.c_int, _ip, C.c_void_p]
    L.amx_target_host_batched.restype = C.c_void_p
    L.amx_target_host_batched.argtypes = [C.c_int, _ip, C.c_void_p, C.c_void_p]
    L.amx_target_destroy.argtypes = [C.c_void_p]
    L.amx_target_eval.argtypes = [C.c_void_p, C.c_long, _ip, _dp, C.c_long, _dp]
    L.amx_proposal_create.restype = C.c_void_p
    L.amx_proposal_create.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, _dp]
    L.amx_proposal_destroy.argtypes = [C.c_void_p]
    L.amx_rj_create.restype = C.c_void_p
    L.amx_rj_create.argtypes = [C.c_void_p, C.c_void_p, C.c_long, _dp, C.c_uint64, C.c_int]
    L.amx_rj_destroy.argtypes = [C.c_void_p]
    L.amx_rj_set_tape.argtypes = [C.c_void_p, _dp, C.c_long]
    L.amx_rj_set_chain_base.argtypes = [C.c_void_p, C.c_uint64]
    L.amx_rj_set_pk_mode.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.amx_rj_set_sort.argtypes = [C.c_void_p, C.c_int]
    L.amx_rj_get_pk_shared.argtypes = [C.c_void_p, _dp, _ip, _dp]
    L.amx_rj_set_modes.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.amx_rwm_set_dof.argtypes = [C.c_int]
    L.amx_copy_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.amx_rj_init_chains.argtypes = [C.c_void_p]
    L.amx_rj_set_state.argtypes = [C.c_void_p, C.c_long, C.c_long, _dp, _dp, _dp, _ip, _ip, _dp, C.c_ulonglong]
    L.amx_rj_get_state.argtypes = [C.c_void_p, C.c_long, C.c_long, _dp, _dp, _dp, _ip, _ip, _dp, _u64p]
    L.amx_rj_sweeps.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int]
    L.amx_rj_collect.argtypes = [C.c_void_p, _u64p, C.POINTER(RjStats), C.c_int]
    L.amx_rj_get_trace.argtypes = [C.c_void_p, _ip, _dp, _dp, _dp]
    L.amx_rj_visit_se.argtypes = [C.c_void_p, _dp, _dp, _ip]
    L.amx_sokal.argtypes = [C.c_int, C.c_long, _dp, _dp, _dp, _ip]
    L.amx_sokal_dev.argtypes = [C.c_int, C.c_long, C.c_void_p, _dp, _dp, _ip]
    L.amx_rj_sokal.argtypes = [C.c_void_p, C.c_long, _dp, _dp, _ip]
    L.amx_rj_moments_reset.argtypes = [C.c_void_p]
    L.amx_rj_moments_accumulate.argtypes = [C.c_void_p]
    L.amx_rj_moments_get.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_ulonglong), _dp, _dp, _dp]
    L.amx_rj_visits_dev.restype = C.c_void_p
    L.amx_rj_visits_dev.argtypes = [C.c_void_p]
    if hasattr(L, "amx_em_fit"):
        L.amx_em_fit.argtypes = [C.c_int, C.c_long, _dp, C.c_int, C.c_int, _ip, _dp, _dp, _dp, _ip, _dp, _dp,
                                 _ip, _dp, _dp, _dp, _ip, _dp, C.POINTER(EmResult)]
        L.amx_em_fit_dev.argtypes = [C.c_int, C.c_long, C.c_void_p, C.c_int, C.c_int, _ip, _dp, _dp, _dp, _ip,
                                     _dp, _dp, _ip, C.POINTER(EmResult)]
        L.amx_em_fit_multi.argtypes = [C.c_int, _ip, C.c_int, C.c_long, _dp, C.c_int, C.c_int, _ip, _dp, _dp, _dp, _ip,
                                       _dp, _dp, _ip, _dp, _dp, _dp, _ip, _dp, C.POINTER(EmResult)]
        L.amx_em_draw_init.restype = C.c_long
        L.amx_em_draw_init.argtypes = [C.c_long, C.c_int, _dp, C.c_long, _ip]
        L.amx_autorj_fit.argtypes = [C.c_int, C.c_long, _dp, _dp, _dp, _dp]
    if hasattr(L, "amx_rwm_adapt"):
        L.amx_rwm_adapt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_long, _dp, C.c_uint64, _dp, C.c_long,
                                    _dp, _dp, _dp, _dp, _dp]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise AmxError(f"amx error {rc}: {lib().amx_last_error().decode()}")


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def device_count() -> int:
    return int(lib().amx_device_count())


def set_stream(ptr: int | None):
    check(lib().amx_set_stream(C.c_void_p(ptr or 0)))


def synchronize():
    check(lib().amx_synchronize())


def set_deferred_sync(on: bool):
    """amx_set_deferred_sync: state transfers only enqueue; buffers must be pinned; sync before touching them."""
    check(lib().amx_set_deferred_sync(int(bool(on))))


def launch_count(reset=False) -> int:
    return int(lib().amx_launch_count(int(reset)))


def measure_fp64_peak() -> float:
    v = C.c_double(0)
    check(lib().amx_measure_fp64_peak(C.byref(v)))
    return v.value


def mix_logpdf(wt, mean, tri, x, want_comp=True, want_mix=True):
    wt, mean, tri, x = f64(wt), f64(mean), f64(tri), f64(x)
    n, d = x.shape
    L = len(wt)
    comp = np.zeros((n, L)) if want_comp else None
    mix = np.zeros(n) if want_mix else None
    check(lib().amx_mix_logpdf(d, L, _d(wt), _d(mean), _d(tri), n, _d(x), _d(comp), _d(mix)))
    return comp, mix


class Target:
    """A log-posterior plug-in handle built from a workload target spec."""

    def __init__(self, spec, host_fn=None, host_batched=None):
        L = lib()
        self.spec = spec
        self.dims = i32(spec["dims"])
        kind = spec["kind"]
        if host_fn is not None:
            self.h = L.amx_target_host_scalar(len(self.dims), _i(self.dims), host_fn)
        elif host_batched is not None:
            self.h = L.amx_target_host_batched(len(self.dims), _i(self.dims), host_batched, None)
        elif kind == "gaussmix":
            self.h = L.amx_target_gaussmix(len(self.dims), _i(self.dims), _i(i32(spec["ncomp"])),
                                           _d(f64(spec["modw"])), _d(f64(spec["wt"])), _d(f64(spec["mean"])),
                                           _d(f64(spec["tri"])), int(spec.get("flags", 0)))
        elif kind == "quad":
            lo, hi = spec.get("lo"), spec.get("hi")
            self.h = L.amx_target_quad(len(self.dims), _i(self.dims), _d(f64(spec["center"])),
                                       _d(f64(spec["scale"])), _d(f64(lo)) if lo is not None else None,
                                       _d(f64(hi)) if hi is not None else None)
        elif kind == "coalmine":
            self.h = L.amx_target_coalmine()
        elif kind == "plugin":  # {"kind": "plugin", "so": path, "dims": ..., "blob": bytes-like, "flags": int}
            blob = np.frombuffer(bytes(spec.get("blob", b"")), dtype=np.uint8)
            self._blob = blob
            self.h = L.amx_target_plugin(str(spec["so"]).encode(), len(self.dims), _i(self.dims),
                                         blob.ctypes.data_as(C.c_void_p) if len(blob) else None, len(blob), int(spec.get("flags", 0)))
        elif kind == "mixnorm":
            self.h = L.amx_target_mixnorm(len(self.dims), _i(i32(spec["ncomp"])), len(spec["y"]), _d(f64(spec["y"])),
                                          _d(f64(spec["prior"])))
        else:
            raise ValueError(kind)
        if not self.h:
            raise AmxError(L.amx_last_error().decode())

    def eval(self, k, x):
        k = i32(k)
        x = f64(x)
        n, ldx = x.shape
        out = np.zeros(n)
        check(lib().amx_target_eval(self.h, n, _i(k), _d(x), ldx, _d(out)))
        return out

    def close(self):
        if self.h:
            lib().amx_target_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Proposal:
    def __init__(self, mix):
        L = lib()
        self.mix = mix
        self.dims = i32(mix["dims"])
        self.h = L.amx_proposal_create(len(self.dims), _i(self.dims), _i(i32(mix["ncomp"])), _d(f64(mix["wt"])),
                                       _d(f64(mix["mean"])), _d(f64(mix["tri"])), _d(f64(mix["sig"])))
        if not self.h:
            raise AmxError(L.amx_last_error().decode())

    def close(self):
        if self.h:
            lib().amx_proposal_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RjPopulation:
    """A population of independent reversible-jump chains on one GPU (amx_rj_*)."""

    def __init__(self, proposal: Proposal, target: Target, nchains: int, init_flat, seed=0, n_trace=0):
        L = lib()
        self.proposal, self.target = proposal, target
        self.C = int(nchains)
        self.nm = len(proposal.dims)
        self.dmax = int(proposal.dims.max())
        self.n_trace = min(int(n_trace), self.C)
        self.h = L.amx_rj_create(proposal.h, target.h, self.C, _d(f64(init_flat)), seed, self.n_trace)
        if not self.h:
            raise AmxError(L.amx_last_error().decode())
        self.last_nsweeps = 0

    def set_modes(self, dof=0, do_perm=False):
        check(lib().amx_rj_set_modes(self.h, int(dof), int(do_perm)))

    def set_pk_mode(self, population: bool, segment_sweeps: int = 0):
        """amx_rj_set_pk_mode: per-chain pk adaptation (the reference's rule) or one pk shared by the population."""
        check(lib().amx_rj_set_pk_mode(self.h, 1 if population else 0, int(segment_sweeps)))

    def set_sort(self, sweeps_per_sort: int = -1):
        """amx_rj_set_sort: re-sort the chains by (model, proposed model) before every n sweeps (-1 automatic, 0 off)."""
        check(lib().amx_rj_set_sort(self.h, int(sweeps_per_sort)))

    def pk_shared(self):
        pk = np.zeros(self.nm)
        nre, lim = C.c_int(0), C.c_double(0)
        check(lib().amx_rj_get_pk_shared(self.h, _d(pk), C.byref(nre), C.byref(lim)))
        return pk, int(nre.value), float(lim.value)

    def set_chain_base(self, first_chain_id: int):
        check(lib().amx_rj_set_chain_base(self.h, int(first_chain_id)))

    def visits_to(self, dst_dev_ptr: int):
        """device-to-device copy of the 64-bit visit histogram (for an NCCL all-reduce)"""
        check(lib().amx_copy_dev(C.c_void_p(dst_dev_ptr), C.c_void_p(self.visits_dev_ptr()), 8 * self.nm))

    def set_tape(self, tape):
        tape = f64(tape)
        assert tape.shape[0] == self.C
        check(lib().amx_rj_set_tape(self.h, _d(tape), tape.shape[1]))

    def init_chains(self):
        check(lib().amx_rj_init_chains(self.h))

    def set_state(self, states, sweep_i, first=0):
        n = len(states)
        theta = np.zeros((n, self.dmax))
        pk = np.zeros((n, self.nm))
        lp = np.zeros(n)
        k = np.zeros(n, np.int32)
        nre = np.zeros(n, np.int32)
        lim = np.zeros(n)
        for i, s in enumerate(states):
            theta[i, : len(s["theta"])] = s["theta"]
            pk[i] = s["pk"]
            lp[i], k[i], nre[i], lim[i] = s["lp"], s["k"], s["nreinit"], s["pkllim"]
        check(lib().amx_rj_set_state(self.h, first, n, _d(theta), _d(pk), _d(lp), _i(k), _i(nre), _d(lim), sweep_i))

    def set_state_arrays(self, st, sweep_i, first=0):
        """Upload chain states from chain-major host arrays (a dict as returned by get_state)."""
        n = len(st["lp"])
        check(lib().amx_rj_set_state(self.h, first, n, _d(st["theta"]), _d(st["pk"]), _d(st["lp"]), _i(st["k"]),
                                     _i(st["nreinit"]), _d(st["pkllim"]), sweep_i))

    def get_state(self, first=0, count=None, out=None):
        n = self.C - first if count is None else count
        if out is not None:
            theta, pk, lp, k, nre, lim = (out[q] for q in ("theta", "pk", "lp", "k", "nreinit", "pkllim"))
        else:
            theta = np.zeros((n, self.dmax))
            pk = np.zeros((n, self.nm))
            lp = np.zeros(n)
            k = np.zeros(n, np.int32)
            nre = np.zeros(n, np.int32)
            lim = np.zeros(n)
        sw = C.c_ulonglong(0)
        check(lib().amx_rj_get_state(self.h, first, n, _d(theta), _d(pk), _d(lp), _i(k), _i(nre), _d(lim), C.byref(sw)))
        return dict(theta=theta, pk=pk, lp=lp, k=k, nreinit=nre, pkllim=lim, sweep_i=int(sw.value))

    def sweeps(self, nsweeps, burning=False, do_adapt=True):
        check(lib().amx_rj_sweeps(self.h, int(nsweeps), int(burning), int(do_adapt)))
        self.last_nsweeps = int(nsweeps)

    def collect(self, reset=False):
        vis = np.zeros(self.nm, np.uint64)
        st = RjStats()
        check(lib().amx_rj_collect(self.h, vis.ctypes.data_as(_u64p), C.byref(st), int(reset)))
        stats = {f: getattr(st, f) for f, _ in RjStats._fields_}
        return vis, stats

    def visit_se(self):
        """(p, se, ngroups) of the counts of the last collect(): standard error from the spread between groups of chains."""
        p, se, g = np.zeros(self.nm), np.zeros(self.nm), C.c_int(0)
        check(lib().amx_rj_visit_se(self.h, _d(p), _d(se), C.byref(g)))
        return p, se, int(g.value)

    def trace(self):
        n = self.last_nsweeps
        k = np.zeros((self.n_trace, n), np.int32)
        lp = np.zeros((self.n_trace, n))
        th = np.zeros((self.n_trace, n, self.dmax))
        pk = np.zeros((self.n_trace, n, self.nm))
        check(lib().amx_rj_get_trace(self.h, _i(k), _d(lp), _d(th), _d(pk)))
        return dict(k=k, lp=lp, theta=th, pk=pk)

    def visits_dev_ptr(self) -> int:
        return int(lib().amx_rj_visits_dev(self.h))

    # -- posterior summaries on the device ----------------------------------------------------
    def sokal(self, nkeep: int):
        """(var, tau, m) per trace chain over the last ``nkeep`` sweeps of the last sweeps call."""
        var, tau, m = np.zeros(self.n_trace), np.zeros(self.n_trace), np.zeros(self.n_trace, np.int32)
        check(lib().amx_rj_sokal(self.h, int(nkeep), _d(var), _d(tau), _i(m)))
        return var, tau, m

    def moments_reset(self):
        check(lib().amx_rj_moments_reset(self.h))

    def moments_accumulate(self):
        check(lib().amx_rj_moments_accumulate(self.h))

    def moments(self, model: int, d: int):
        cnt = C.c_ulonglong()
        mean, cov, mlp = np.zeros(d), np.zeros((d, d)), C.c_double()
        check(lib().amx_rj_moments_get(self.h, int(model), C.byref(cnt), _d(mean), _d(cov), C.byref(mlp)))
        return dict(count=int(cnt.value), mean=mean, cov=cov, mean_lp=mlp.value)

    def close(self):
        if self.h:
            lib().amx_rj_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FamHdr(C.Structure):
    """amx_fam_hdr of include/amx_layout.h."""
    _fields_ = [("nmodels", C.c_int), ("dmax", C.c_int), ("Lmax", C.c_int), ("total", C.c_int)] + [
        (n, C.c_int * 32) for n in ("dims", "ncomp", "off", "stride", "ext", "extlen")]


def family_blob(spec) -> bytes:
    """The device family blob [amx_fam_hdr | doubles] of a Gaussian-mixture target spec (amx_fam_plan + amx_fam_pack with
    AMX_FAM_TARGET): the parameter block a user plug-in for such a target receives in bind()."""
    L = lib()
    L.amx_fam_plan.argtypes = [C.POINTER(FamHdr), C.c_int, _ip, _ip, _ip]
    L.amx_fam_pack.argtypes = [C.POINTER(FamHdr), C.c_int, _dp, _dp, _dp, _dp, _dp]
    h = FamHdr()
    dims, ncomp = i32(spec["dims"]), i32(spec["ncomp"])
    ext = np.ones(len(dims), np.int32)
    total = L.amx_fam_plan(C.byref(h), len(dims), _i(dims), _i(ncomp), _i(ext))
    if total < 0:
        raise AmxError("bad family shape")
    data = np.zeros(total)
    L.amx_fam_pack(C.byref(h), 1, _d(f64(spec["wt"])), _d(f64(spec["mean"])), _d(f64(spec["tri"])), _d(f64(spec["modw"])), _d(data))
    return bytes(h) + data.tobytes()


def em_draw_init(n, Lmax, uniforms):
    u = f64(uniforms)
    idx = np.zeros(Lmax, np.int32)
    used = lib().amx_em_draw_init(n, Lmax, _d(u), len(u), _i(idx))
    if used < 0:
        raise AmxError("uniform stream exhausted while drawing EM start rows")
    return idx, int(used)


def em_fit(x, init_idx, Lmax=30, maxit=5000, want_state=False, x_dev_ptr=None, devices=None):
    """Figueiredo-Jain EM fit on the GPU.  x: (n,d) host array, or pass x_dev_ptr (+ shape via x).
    devices: list of CUDA ordinals to shard the samples over (amx_em_fit_multi)."""
    L = lib()
    n, d = x.shape
    t = d * (d + 1) // 2
    init_idx = i32(init_idx)
    wt = np.zeros(Lmax)
    mean = np.zeros((Lmax, d))
    tri = np.zeros((Lmax, t))
    cap = maxit + 1
    trL = np.zeros(cap, np.int32)
    trll = np.zeros(cap)
    trc = np.zeros(cap)
    tra = np.zeros(cap, np.int32)
    res = EmResult()
    st = {}
    if x_dev_ptr is not None:
        check(L.amx_em_fit_dev(d, n, C.c_void_p(x_dev_ptr), Lmax, maxit, _i(init_idx), _d(wt), _d(mean), _d(tri),
                               _i(trL), _d(trll), _d(trc), _i(tra), C.byref(res)))
    else:
        x = f64(x)
        if want_state:
            st = dict(cur_wt=np.zeros(Lmax), cur_mean=np.zeros((Lmax, d)), cur_tri=np.zeros((Lmax, t)),
                      cur_L=np.zeros(1, np.int32), cur_w=np.zeros((n, Lmax)))
        if devices is not None:
            dv = i32(devices)
            check(L.amx_em_fit_multi(len(dv), _i(dv), d, n, _d(x), Lmax, maxit, _i(init_idx), _d(wt), _d(mean), _d(tri),
                                     _i(trL), _d(trll), _d(trc), _i(tra), _d(st.get("cur_wt")), _d(st.get("cur_mean")),
                                     _d(st.get("cur_tri")), _i(st.get("cur_L")), _d(st.get("cur_w")), C.byref(res)))
        else:
            check(L.amx_em_fit(d, n, _d(x), Lmax, maxit, _i(init_idx), _d(wt), _d(mean), _d(tri), _i(trL), _d(trll),
                               _d(trc), _i(tra), _d(st.get("cur_wt")), _d(st.get("cur_mean")), _d(st.get("cur_tri")),
                               _i(st.get("cur_L")), _d(st.get("cur_w")), C.byref(res)))
    Lb, it = res.L, res.iters
    out = dict(L=Lb, iters=it, lam=wt[:Lb].copy(), mu=mean[:Lb].copy(), B=tri[:Lb].copy(), trace_L=trL[:it].copy(),
               trace_loglik=trll[:it].copy(), trace_cost=trc[:it].copy(), trace_ann=tra[:it].copy(),
               kernel_ms=res.kernel_ms, comp_steps=res.comp_steps, flops=res.flops, bytes=res.bytes,
               bytes_requested=res.bytes_requested, status=res.status)
    if st:
        cl = int(st["cur_L"][0])
        out.update(cur_L=cl, cur_lam=st["cur_wt"][:cl].copy(), cur_mu=st["cur_mean"][:cl].copy(),
                   cur_B=st["cur_tri"][:cl].copy(), cur_w=st["cur_w"][:, :cl].copy())
    return out


def autorj_fit(x):
    x = f64(x)
    n, d = x.shape
    wt = np.zeros(1)
    mean = np.zeros(d)
    tri = np.zeros(d * (d + 1) // 2)
    check(lib().amx_autorj_fit(d, n, _d(x), _d(wt), _d(mean), _d(tri)))
    return dict(lam=wt, mu=mean, B=tri)


def rwm_adapt(target: Target, model_k: int, nsweep2: int, nchains: int, init, seed=0, tapes=None, dof=0):
    """Stage-1 adaptive RWM for one model on the GPU (amx_rwm_adapt)."""
    check(lib().amx_rwm_set_dof(int(dof)))
    d = int(target.dims[model_k])
    init = f64(init)
    nsw = max(nsweep2, 10000 * d)
    total = nsw + nsw // 10
    sig = np.zeros((nchains, d))
    samples = np.zeros((nchains, 1000 * d, d))
    ntr = total // 100
    sig_tr = np.zeros((ntr, d))
    acc_tr = np.zeros((ntr, d))
    ms = C.c_double(0)
    tp, stride = None, 0
    if tapes is not None:
        tapes = f64(tapes)
        assert tapes.shape[0] == nchains
        tp, stride = _d(tapes), tapes.shape[1]
    check(lib().amx_rwm_adapt(target.h, model_k, nsweep2, nchains, _d(init), seed, tp, stride, _d(sig), _d(samples),
                              _d(sig_tr), _d(acc_tr), C.byref(ms)))
    return dict(sweeps=total, sig=sig, samples=samples, sig_trace=sig_tr, acc_trace=acc_tr, kernel_ms=ms.value)
