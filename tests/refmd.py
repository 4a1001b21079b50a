"""Whole Verlet step on the CPU with every function the reference SHIPS run from its own compiled translation units
(oracle/_ref/libref.so: Pair*Intel::compute, PPPMIntel::compute, FixNVEIntel::initial_integrate / final_integrate) and
only what the reference inherits from upstream LAMMPS — Neighbor::decide, Domain::pbc, Comm::borders / forward_comm /
reverse_comm, the binned half list — supplied by the oracle's restatement of it.  TEST INFRASTRUCTURE ONLY: the checker
of tests/test_oracle_vs_ref.py and the CPU arm of bench.py (`--impl reference`, `cpu_baseline`).

The step order is stock Verlet::run, the same as oracle/md.cpp:
    initial_integrate -> decide -> [pbc, borders, build | forward_comm] -> pair -> reverse_comm -> kspace -> final_integrate

Timers: `ref_*` = wall time inside the reference's own member functions (ref_last_seconds: harness set-up excluded),
`glue_*` = the upstream pieces, `harness` = marshalling between the two libraries (object set-up per call, copies) —
something a LAMMPS build does not do and the reported throughput therefore leaves out."""
import time

import numpy as np

import orc
import refc


class RefMD:
    def __init__(self, system, params, pppm=None, prec=orc.DOUBLE, skin=0.3, every=1, delay=0, check=1, dt=0.001,
                 ftm2v=1.0, nthreads=1):
        s = system
        self.n = len(s["x"])
        self.x = orc.f64(s["x"]).copy()
        self.v = orc.f64(s["v"]).copy() if s.get("v") is not None else np.zeros((self.n, 3))
        self.q = orc.f64(s["q"]).copy() if s.get("q") is not None else np.zeros(self.n)
        self.type = orc.i32(s["type"]).copy()
        self.mass = orc.f64(s["mass"])
        self.ntypes = int(s["ntypes"])
        self.lo, self.hi = orc.f64(s["boxlo"]).copy(), orc.f64(s["boxhi"]).copy()
        self.prd = self.hi - self.lo
        self.P, self.pp, self.prec = params, pppm, prec
        self.skin, self.every, self.delay, self.check = skin, every, delay, check
        self.dt, self.ftm2v, self.nthreads = dt, ftm2v, nthreads
        self.f = np.zeros((self.n, 3))
        self.ago, self.built, self.nbuilds = 0, False, 0
        self.t = dict(ref_pair=0.0, ref_kspace=0.0, ref_nve=0.0, glue_neigh=0.0, glue_comm=0.0, harness=0.0)

    # ---- upstream glue (oracle restatement) ----------------------------------------------------------------------
    def _pbc(self):
        for d in range(3):
            c = self.x[:, d]
            c[c < self.lo[d]] += self.prd[d]
            hi = c >= self.hi[d]
            c[hi] = np.maximum(c[hi] - self.prd[d], self.lo[d])

    def _build(self):
        t0 = time.perf_counter()
        self._pbc()
        cm = self.P.cutmax() + self.skin
        self.xa, self.ta, self.qa, self.src, self.shift = orc.make_ghosts(self.x, self.type, self.q, self.lo, self.hi, cm)
        self.nn, self.off, self.ent = orc.neigh_half_bin(self.n, self.xa, self.ta, self.ntypes, self.P.cutneighsq(self.skin),
                                                         self.lo, self.hi, cm, self.prec)
        # the owned atom at the root of every ghost's source chain (ghosts are made dimension by dimension, so a ghost
        # may copy a ghost); `shift` is the total image shift from that root
        root = self.src.copy()
        for _ in range(3):
            g = root >= self.n
            if not g.any():
                break
            root[g] = self.src[root[g] - self.n]
        self.root = root
        self.xhold = self.x.copy()
        self.ago, self.built = 0, True
        self.nbuilds += 1
        self.t["glue_neigh"] += time.perf_counter() - t0

    def _forward_comm(self):
        t0 = time.perf_counter()
        self.xa[:self.n] = self.x
        self.xa[self.n:] = self.x[self.root] + self.shift * self.prd
        self.t["glue_comm"] += time.perf_counter() - t0

    # ---- the reference's members -------------------------------------------------------------------------------------
    def _forces(self):
        t0 = time.perf_counter()
        ref = 0.0
        f4, _ = refc.pair_eval(self.P, self.prec, 0, 0, self.n, self.xa, self.ta, self.qa, self.nn, self.off, self.ent,
                               newton=1, nthreads=self.nthreads, skin=self.skin)
        tp = refc.last_seconds()
        self.t["ref_pair"] += tp
        t1 = time.perf_counter()
        f4 = orc.reverse_comm(self.n, self.src, f4)
        self.t["glue_comm"] += time.perf_counter() - t1
        glue = time.perf_counter() - t1
        self.f = np.ascontiguousarray(f4[:self.n, :3])
        tk = 0.0
        if self.pp is not None:
            fk = refc.pppm_compute(self.pp, self.x, self.q, prec=self.prec, eflag=0, vflag=0, nthreads=self.nthreads,
                                   want_grids=False)[0]
            tk = refc.last_seconds()
            self.t["ref_kspace"] += tk
            self.f += fk
        ref = tp + tk
        self.t["harness"] += time.perf_counter() - t0 - ref - glue

    def _nve(self, which):
        t0 = time.perf_counter()
        self.x, self.v = refc.nve(which, self.x, self.v, self.f, self.type, self.mass, self.dt, self.ftm2v)
        tr = refc.last_seconds()
        self.t["ref_nve"] += tr
        self.t["harness"] += time.perf_counter() - t0 - tr

    def run(self, nsteps):
        if not self.built:
            self._build()
            self._forces()
        for _ in range(nsteps):
            self._nve(0)
            self.ago += 1
            rebuild = False
            if self.ago >= self.delay and self.ago % self.every == 0:
                if self.check:
                    t0 = time.perf_counter()
                    d = self.x - self.xhold
                    rebuild = bool((np.einsum("ij,ij->i", d, d) > 0.25 * self.skin * self.skin).any())
                    self.t["glue_neigh"] += time.perf_counter() - t0
                else:
                    rebuild = True
            if rebuild:
                self._build()
            else:
                self._forward_comm()
            self._forces()
            self._nve(1)
        return dict(self.t, nbuilds=self.nbuilds)

    def seconds(self):
        """time a LAMMPS build would spend: the reference's members + the upstream glue (marshalling excluded)"""
        return sum(v for k, v in self.t.items() if k != "harness")
