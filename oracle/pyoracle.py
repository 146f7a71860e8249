"""ctypes doorway to the parity checkers.  TEST INFRASTRUCTURE ONLY.

Two libraries expose the same flat API under two prefixes:

* ``orc_*``  -- oracle/_build/libamx_oracle.so, our CPU restatement (amx_oracle.c)
* ``ref_*``  -- oracle/_ref/libautomix_tape.so, the UNMODIFIED reference compiled from
  /root/reference (ref_harness.c), present wherever ``make -C oracle ref`` has been run
  (this container; it then travels to the GPU box with the snapshot).

Only tests/, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of bench.py may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libamx_oracle.so")
TARGETS_SO = os.path.join(HERE, "_build", "libamx_hosttargets.so")
REF_TAPE_SO = os.path.join(HERE, "_ref", "libautomix_tape.so")
REF_SO = os.path.join(HERE, "_ref", "libautomix.so")
REF_USERTARGETS_SO = os.path.join(HERE, "_ref", "libref_usertargets.so")
REF_LOGWRITE_SO = os.path.join(HERE, "_ref", "libref_logwrite.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_ulp = C.POINTER(C.c_ulong)
_lp = C.POINTER(C.c_long)
TARGET_FN = C.CFUNCTYPE(C.c_double, C.c_int, _dp)


def build(ref: bool = True) -> None:
    """(Re)build the checkers with oracle/Makefile."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/src/libautomix"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def have_ref() -> bool:
    return os.path.exists(REF_TAPE_SO)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def tri_len(d):
    return d * (d + 1) // 2


def pack_lower(M):
    """dense lower-triangular (d,d) -> packed row-major."""
    M = np.asarray(M, dtype=np.float64)
    d = M.shape[0]
    return np.array([M[i, j] for i in range(d) for j in range(i + 1)], dtype=np.float64)


def unpack_lower(p, d):
    M = np.zeros((d, d))
    t = 0
    for i in range(d):
        for j in range(i + 1):
            M[i, j] = p[t]
            t += 1
    return M


class Checker:
    """One of the two checker libraries (``prefix`` = 'orc' or 'ref')."""

    def __init__(self, prefix: str):
        assert prefix in ("orc", "ref")
        self.prefix = prefix
        path = ORACLE_SO if prefix == "orc" else REF_TAPE_SO
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self._tape = None
        g = self._fn
        g("tape_set", None, [_dp, C.c_long])
        g("tape_used", C.c_long, [])
        g("tape_overrun", C.c_int, [])
        g("gauss", None, [_dp, C.c_int])
        g("rt", None, [_dp, C.c_int, C.c_int])
        g("perm", None, [_dp, C.c_int])
        g("rgamma", C.c_double, [C.c_double])
        g("loggamma", C.c_double, [C.c_double])
        g("ltprob", C.c_double, [C.c_int, C.c_double])
        g("chol", None, [C.c_int, _dp])
        g("det", C.c_double, [C.c_int, _dp])
        g("lnormprob", C.c_double, [C.c_int, _dp, _dp, _dp])
        g("mix_logpdf", None, [C.c_int, C.c_int, _dp, _dp, _dp, C.c_long, _dp, _dp, _dp])
        g("rwm_within_model", C.c_int,
          [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp, _dp])
        g("fit_mixture", C.c_int,
          [C.c_int, C.c_int, _dp, C.c_int, C.c_int, _dp, _dp, _dp, _ip, _ip, _dp, _dp, _ip, _ip,
           _dp, _dp, _dp, _ip, _dp, _dp])
        g("fit_autorj", None, [C.c_int, C.c_int, _dp, _dp, _dp, _dp])
        g("chain_init", C.c_int,
          [C.c_int, _ip, _dp, C.c_void_p, _dp, _dp, _dp, _ip, _ip, _dp, _ulp])
        g("rj_sweeps", C.c_int,
          [C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_int,
           C.c_int, _dp, _dp, _dp, _ip, _ip, _dp, _ulp, _ip, _dp, _dp, _dp, _ulp, _lp])

        if prefix == "orc":  # posterior summaries: the reference side is RefLogwrite.sokal
            g("sokal", C.c_int, [C.c_long, _dp, _dp, _dp, _ip])
            g("model_moments", C.c_long, [C.c_long, C.c_int, _ip, _dp, C.c_int, C.c_int, _dp, _dp])

    def sokal(self, x):
        """(var, tau, m) of one series, logwrite.c:354-403."""
        x = f64(x)
        var, tau, m = C.c_double(), C.c_double(), C.c_int()
        rc = self._sokal(len(x), _d(x), C.byref(var), C.byref(tau), C.byref(m))
        if rc != 0:
            raise ValueError("sokal: length must be a power of two in [4, 2^20]")
        return var.value, tau.value, m.value

    def model_moments(self, k, theta, model, d):
        k = i32(k)
        theta = f64(theta)
        mean, cov = np.zeros(d), np.zeros((d, d))
        cnt = self._model_moments(len(k), theta.shape[1], _i(k), _d(theta), model, d, _d(mean), _d(cov))
        return int(cnt), mean, cov

    def _fn(self, name, res, args):
        f = getattr(self.lib, f"{self.prefix}_{name}")
        f.restype = res
        f.argtypes = args
        setattr(self, "_" + name, f)

    # -- uniform tape ---------------------------------------------------------
    def tape(self, u):
        self._tape = f64(u)  # keep alive
        self._tape_set(_d(self._tape), len(self._tape))

    def tape_used(self):
        return int(self._tape_used())

    def tape_overrun(self):
        return bool(self._tape_overrun())

    # -- helpers ---------------------------------------------------------------
    def gauss(self, n):
        z = np.zeros(n)
        self._gauss(_d(z), n)
        return z

    def rt(self, n, dof):
        z = np.zeros(n)
        self._rt(_d(z), n, dof)
        return z

    def perm(self, v):
        v = f64(v).copy()
        self._perm(_d(v), len(v))
        return v

    def rgamma(self, s):
        return float(self._rgamma(s))

    def loggamma(self, x):
        return float(self._loggamma(x))

    def ltprob(self, dof, z):
        return float(self._ltprob(dof, z))

    def chol(self, packed, d):
        a = f64(packed).copy()
        self._chol(d, _d(a))
        return a

    def det(self, packed, d):
        return float(self._det(d, _d(f64(packed))))

    def lnormprob(self, mu, Bp, x):
        mu, Bp, x = f64(mu), f64(Bp), f64(x)
        return float(self._lnormprob(len(mu), _d(mu), _d(Bp), _d(x)))

    def mix_logpdf(self, wt, mean, tri, x):
        wt, mean, tri, x = f64(wt), f64(mean), f64(tri), f64(x)
        L = len(wt)
        n, d = x.shape
        comp = np.zeros((n, L))
        mix = np.zeros(n)
        self._mix_logpdf(d, L, _d(wt), _d(mean), _d(tri), n, _d(x), _d(comp), _d(mix))
        return comp, mix

    # -- stage 1 ---------------------------------------------------------------
    def rwm_within_model(self, model_k, d, nsweep2, target_ptr, init, dof=0):
        nsw = max(nsweep2, 10000 * d)
        total = nsw + nsw // 10
        sig = np.zeros(d)
        samples = np.zeros((1000 * d, d))
        ntr = total // 100
        sig_tr = np.zeros((ntr, d))
        acc_tr = np.zeros((ntr, d))
        fin = np.zeros(d)
        flp = np.zeros(1)
        init = f64(init)
        n = self._rwm_within_model(model_k, d, nsweep2, dof, target_ptr, _d(init), _d(sig),
                                   _d(samples), _d(sig_tr), _d(acc_tr), _d(fin), _d(flp))
        return dict(sweeps=n, sig=sig, samples=samples, sig_trace=sig_tr, acc_trace=acc_tr,
                    final=fin, final_lp=float(flp[0]))

    # -- stage 2 ---------------------------------------------------------------
    def fit_mixture(self, x, Lmax=30, maxit=5000, want_state=False):
        x = f64(x)
        n, d = x.shape
        t = tri_len(d)
        lam = np.zeros(Lmax)
        mu = np.zeros((Lmax, d))
        B = np.zeros((Lmax, t))
        Lout = np.zeros(1, np.int32)
        cap = maxit + 2
        trL = np.zeros(cap, np.int32)
        trll = np.zeros(cap)
        trc = np.zeros(cap)
        tra = np.zeros(cap, np.int32)
        idx = np.zeros(Lmax, np.int32)
        st = {}
        if want_state and self.prefix == "orc":
            st = dict(cur_lam=np.zeros(Lmax), cur_mu=np.zeros((Lmax, d)), cur_B=np.zeros((Lmax, t)),
                      cur_L=np.zeros(1, np.int32), cur_w=np.zeros((n, Lmax)),
                      cur_lpd=np.zeros((n, Lmax)))
        it = self._fit_mixture(d, n, _d(x), Lmax, maxit, _d(lam), _d(mu), _d(B), _i(Lout), _i(trL),
                               _d(trll), _d(trc), _i(tra), _i(idx), _d(st.get("cur_lam")),
                               _d(st.get("cur_mu")), _d(st.get("cur_B")), _i(st.get("cur_L")),
                               _d(st.get("cur_w")), _d(st.get("cur_lpd")))
        L = int(Lout[0])
        out = dict(L=L, iters=it, lam=lam[:L].copy(), mu=mu[:L].copy(), B=B[:L].copy(),
                   trace_L=trL[:it].copy(), trace_loglik=trll[:it].copy(), trace_cost=trc[:it].copy(),
                   trace_ann=tra[:it].copy(), init_idx=idx.copy())
        if st:
            cl = int(st["cur_L"][0])
            out.update(cur_L=cl, cur_lam=st["cur_lam"][:cl].copy(), cur_mu=st["cur_mu"][:cl].copy(),
                       cur_B=st["cur_B"][:cl].copy(), cur_w=st["cur_w"][:, :cl].copy(),
                       cur_lpd=st["cur_lpd"][:, :cl].copy())
        return out

    def fit_autorj(self, x):
        x = f64(x)
        n, d = x.shape
        lam = np.zeros(1)
        mu = np.zeros(d)
        B = np.zeros(tri_len(d))
        self._fit_autorj(d, n, _d(x), _d(lam), _d(mu), _d(B))
        return dict(lam=lam, mu=mu, B=B)

    # -- stage 3 ---------------------------------------------------------------
    def chain_init(self, dims, init_flat, target_ptr):
        dims = i32(dims)
        nm = len(dims)
        dmax = int(dims.max())
        theta = np.zeros(dmax)
        pk = np.zeros(nm)
        lp = np.zeros(1)
        k = np.zeros(1, np.int32)
        nre = np.zeros(1, np.int32)
        lim = np.zeros(1)
        sw = np.zeros(1, np.uint64)
        init_flat = f64(init_flat)
        self._chain_init(nm, _i(dims), _d(init_flat), target_ptr, _d(theta), _d(pk), _d(lp), _i(k),
                         _i(nre), _d(lim), sw.ctypes.data_as(_ulp))
        return dict(theta=theta, pk=pk, lp=float(lp[0]), k=int(k[0]), nreinit=int(nre[0]),
                    pkllim=float(lim[0]), sweep_i=int(sw[0]))

    def rj_sweeps(self, mix, target_ptr, state, nsweeps, burning=False, do_adapt=True,
                  do_perm=False, dof=0, trace=True):
        """mix: dict(dims, ncomp, wt, mean, tri, sig) flat; state: dict from chain_init."""
        dims, ncomp = i32(mix["dims"]), i32(mix["ncomp"])
        nm = len(dims)
        dmax = int(dims.max())
        wt, mean, tri, sig = f64(mix["wt"]), f64(mix["mean"]), f64(mix["tri"]), f64(mix["sig"])
        theta = np.zeros(dmax)
        theta[: len(state["theta"])] = state["theta"]
        pk = f64(state["pk"]).copy()
        lp = np.array([state["lp"]])
        k = np.array([state["k"]], np.int32)
        nre = np.array([state["nreinit"]], np.int32)
        lim = np.array([state["pkllim"]])
        sw = np.array([state["sweep_i"]], np.uint64)
        trk = np.zeros(nsweeps, np.int32) if trace else None
        trlp = np.zeros(nsweeps) if trace else None
        trth = np.zeros((nsweeps, dmax)) if trace else None
        trpk = np.zeros((nsweeps, nm)) if trace else None
        cnt = np.zeros(6, np.uint64)
        vis = np.zeros(nm, np.int64)
        rc = self._rj_sweeps(nm, _i(dims), _i(ncomp), _d(wt), _d(mean), _d(tri), _d(sig), target_ptr,
                             nsweeps, int(burning), int(do_adapt), int(do_perm), dof, _d(theta),
                             _d(pk), _d(lp), _i(k), _i(nre), _d(lim), sw.ctypes.data_as(_ulp),
                             _i(trk), _d(trlp), _d(trth), _d(trpk), cnt.ctypes.data_as(_ulp),
                             vis.ctypes.data_as(_lp))
        assert rc == 0
        new = dict(theta=theta, pk=pk, lp=float(lp[0]), k=int(k[0]), nreinit=int(nre[0]),
                   pkllim=float(lim[0]), sweep_i=int(sw[0]))
        return dict(state=new, k=trk, lp=trlp, theta=trth, pk=trpk, counters=cnt, visits=vis)


class HostTargets:
    """Host log-posterior callbacks (oracle/host_targets.c)."""

    def __init__(self):
        self.lib = C.CDLL(TARGETS_SO)
        L = self.lib
        L.amxh_select_gaussmix.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, C.c_int]
        L.amxh_select_quad.argtypes = [C.c_int, _ip, _dp, _dp, _dp, _dp]
        L.amxh_select_mixnorm.argtypes = [C.c_int, _ip, C.c_int, _dp, _dp]
        L.amxh_logpost.restype = C.c_double
        L.amxh_logpost.argtypes = [C.c_int, _dp]
        L.amxh_logpost_ptr.restype = C.c_void_p
        L.amxh_batched_ptr.restype = C.c_void_p
        L.amxh_calls.restype = C.c_long
        L.amxh_calls.argtypes = [C.c_int]
        self.ptr = C.c_void_p(L.amxh_logpost_ptr())
        self.batched_ptr = C.c_void_p(L.amxh_batched_ptr())

    def select(self, spec):
        """spec: a workload target dict (automix_b200.workloads)."""
        kind = spec["kind"]
        if kind == "gaussmix":
            rc = self.lib.amxh_select_gaussmix(
                len(spec["dims"]), _i(i32(spec["dims"])), _i(i32(spec["ncomp"])), _d(f64(spec["modw"])),
                _d(f64(spec["wt"])), _d(f64(spec["mean"])), _d(f64(spec["tri"])), int(spec.get("flags", 0)))
        elif kind == "quad":
            lo = spec.get("lo")
            hi = spec.get("hi")
            rc = self.lib.amxh_select_quad(
                len(spec["dims"]), _i(i32(spec["dims"])), _d(f64(spec["center"])), _d(f64(spec["scale"])),
                _d(f64(lo)) if lo is not None else None, _d(f64(hi)) if hi is not None else None)
        elif kind == "coalmine":
            rc = self.lib.amxh_select_coalmine()
        elif kind == "mixnorm":
            rc = self.lib.amxh_select_mixnorm(len(spec["dims"]), _i(i32(spec["ncomp"])), len(spec["y"]),
                                              _d(f64(spec["y"])), _d(f64(spec["prior"])))
        else:
            raise ValueError(kind)
        assert rc == 0
        return self.ptr

    def logpost(self, k, x):
        x = f64(x)
        return float(self.lib.amxh_logpost(int(k), _d(x)))

    def calls(self, reset=False):
        return int(self.lib.amxh_calls(int(reset)))


class RefUserTargets:
    """The reference's own example log-posteriors (only where oracle/_ref was built)."""

    def __init__(self):
        self.lib = C.CDLL(REF_USERTARGETS_SO)
        for n in ("ref_toy1", "ref_toy2", "ref_cpt"):
            f = getattr(self.lib, n)
            f.restype = C.c_double
            f.argtypes = [C.c_int, _dp]
        self.lib.ref_cpt_init.argtypes = [C.c_int, C.c_int, _dp]

    def eval(self, name, k, x):
        x = f64(x)
        return float(getattr(self.lib, "ref_" + name)(int(k), _d(x)))

    def cpt_init(self, k, d):
        v = np.zeros(d)
        self.lib.ref_cpt_init(k, d, _d(v))
        return v


class RefLogwrite:
    """The reference's report writer compiled as it lies (user_examples/logwrite.c): sokal()."""

    def __init__(self):
        self.lib = C.CDLL(REF_LOGWRITE_SO)
        self.lib.sokal.restype = None
        self.lib.sokal.argtypes = [C.c_int, _dp, _dp, _dp, _ip]

    def sokal(self, x):
        buf = f64(x).copy()  # the reference overwrites its input with the autocorrelations
        var, tau, m = C.c_double(), C.c_double(), C.c_int()
        self.lib.sokal(len(buf), _d(buf), C.byref(var), C.byref(tau), C.byref(m))
        return var.value, tau.value, m.value, buf


def have_ref_logwrite() -> bool:
    return os.path.exists(REF_LOGWRITE_SO)


class RefPristine:
    """The unmodified reference .so with its own generator (pipeline-level runs, CPU timing)."""

    def __init__(self):
        self.lib = C.CDLL(REF_SO)
        self.lib.sdrand.restype = C.c_double
        self.lib.sdrni.argtypes = [_ulp]

    def seed(self, s):
        v = C.c_ulong(s)
        self.lib.sdrni(C.byref(v))

    def uniforms(self, n):
        return np.array([self.lib.sdrand() for _ in range(n)])
