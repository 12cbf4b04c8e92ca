"""Row-block-sharded CLIPPER across the GPUs of one box.

One process per GPU (``torchrun``), ``torch.distributed`` for the plumbing only:
  * the 256-byte peer-memory blobs (CUDA IPC handles) of every rank are all-gathered once,
  * bench timings are max-reduced over ranks.
The data path never goes through NCCL or the host: scoring is embarrassingly parallel (rank r
scores its own row block of M), and the solver is one persistent kernel per GPU that performs
its single exchange step per objective evaluation through NVLink peer memory (P2P stores of the
new gradient entries and of the per-rank partial sums, release/acquire flags).  A collective
formulation of the same exchange (`gathered_matvec`: all-gather of the disjoint slices == all-reduce
of zero-padded vectors) is kept as the cross-check.

`ShardGroup` drives several shards from ONE process (threads): with two GPUs it is the
world-size-2 GPU test, and with `same_device=True` (1 CTA per SM per shard) it exercises the whole
sharded code path on a single GPU.
"""
import ctypes as C
import threading

import numpy as np

from . import _capi
from .api import (CLIPPER, EuclideanDistance, PointNormalDistance, Params, Solution)


def shard_rows(m, rank, world):
    """row block [row0, row0+rows) owned by `rank` (multiples of the 32-row tile)"""
    r0, n = C.c_int64(), C.c_int64()
    _capi.load().clp_shard_rows(int(m), int(rank), int(world), C.byref(r0), C.byref(n))
    return int(r0.value), int(n.value)


def partition_is_exact(m, world):
    """host-side invariant used by the CPU tests: the shards tile [0,m) exactly once, in order"""
    nxt = 0
    for r in range(world):
        r0, n = shard_rows(m, r, world)
        if r0 != nxt or n < 0:
            return False
        nxt = r0 + n
    return nxt == m


def interleave_permutation(m, world):
    """new_of_old[i]: deal the associations to the row shards round-robin (shard sizes as clp_shard_rows cuts them).
    Row-count shards are not byte-balanced when dense rows cluster -- the benchmark generator puts the inlier
    associations last (reference bm_utils.cpp:312-315,344): 1.144x the mean bytes on the last of 8 shards at
    BASELINE config 2.  Scoring the permuted association list permutes M symmetrically; the solution of the
    permuted problem maps back with unpermute_solution().  Host-side only; the default of ShardedCLIPPER's host-pointer
    scoring (balance=True) and of bench.py --gpus N (CLP_SHARD_INTERLEAVE=0 turns it off)."""
    bounds = [shard_rows(m, r, world) for r in range(world)]
    # slot k of a shard with n rows is due at time (k + 0.5) / n: taking all slots in time order fills every shard
    # at a rate proportional to its size, so any run of consecutive associations is spread evenly
    due = np.concatenate([(np.arange(n) + 0.5) / max(n, 1) for _, n in bounds])
    slot = np.concatenate([r0 + np.arange(n) for r0, n in bounds])
    new_of_old = slot[np.argsort(due, kind="stable")].astype(np.int64)
    return new_of_old


def permute_problem(A, u0, new_of_old):
    """association list and start vector in the permuted numbering"""
    A = np.asarray(A); m = A.shape[0]
    Ap = np.empty_like(A); Ap[new_of_old] = A
    u0p = None
    if u0 is not None:
        u0p = np.empty(m, dtype=np.float64); u0p[new_of_old] = np.asarray(u0, dtype=np.float64)
    return np.asfortranarray(Ap), u0p


def unpermute_solution(nodes_new, u_new, new_of_old):
    """node indices and iterate of the permuted problem in the caller's numbering (node order is kept)"""
    old_of_new = np.empty_like(new_of_old); old_of_new[new_of_old] = np.arange(len(new_of_old))
    nodes_old = old_of_new[np.asarray(nodes_new, dtype=np.int64)].astype(np.int32)
    u_old = None if u_new is None else np.asarray(u_new)[new_of_old]
    return nodes_old, u_old


class ShardedCLIPPER(CLIPPER):
    """clipper::CLIPPER whose affinity matrix is row-sharded over the ranks of a process group.
    Every rank calls every method collectively with identical arguments."""

    def __init__(self, invariant, params, group=None, device=None, storage=_capi.STORE_F32):
        import torch
        import torch.distributed as dist
        self._dist = dist
        self._group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if device is None:
            device = torch.cuda.current_device()
        super().__init__(invariant, params, device=device, storage=storage)
        _capi.check(self._h, self._lib.clp_shard_config(self._h, self.rank, self.world))
        self._connected_for = None

    def _connect(self):
        """all-gather the peer-memory blobs (once per buffer generation)"""
        import torch
        nb = int(self._lib.clp_shard_blob_bytes())
        blob = C.create_string_buffer(nb)
        wrote = C.c_int64()
        _capi.check(self._h, self._lib.clp_shard_export(self._h, blob, nb, C.byref(wrote)))
        mine = torch.frombuffer(bytearray(blob.raw), dtype=torch.uint8)
        backend = self._dist.get_backend(self._group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        mine = mine.to(dev)
        allb = torch.empty(self.world * nb, dtype=torch.uint8, device=dev)
        self._dist.all_gather_into_tensor(allb, mine, group=self._group)
        raw = bytes(allb.cpu().numpy().tobytes())
        _capi.check(self._h, self._lib.clp_shard_import(self._h, raw, nb, self.world))
        self._dist.barrier(group=self._group)

    def score_pairwise_consistency(self, D1, D2, A=None, balance=True):
        """balance: relabel the associations so that every row shard holds about the same number of bytes (see
        interleave_permutation); solve() maps nodes / u / u0 back, the caller never sees the relabelling"""
        self._perm = None
        if balance and A is not None and np.size(A) > 0 and self.world > 1:
            self._perm = interleave_permutation(np.asarray(A).shape[0], self.world)
            A, _ = permute_problem(A, None, self._perm)
        super().score_pairwise_consistency(D1, D2, A)
        m = self._m()
        if self._connected_for != m:
            self._connect()
            self._connected_for = m

    def get_initial_associations(self):
        A = super().get_initial_associations()
        perm = getattr(self, "_perm", None)
        return A if perm is None else np.asfortranarray(A[perm])

    def score_device(self, kind, D1, D2, A, *inv_params):
        """device-pointer scoring (torch tensors), same conventions as the C-ABI"""
        L = self._lib
        if kind == "euclidean":
            rc = L.clp_score_euclidean_dev(self._h, D1.data_ptr(), D1.shape[1], D1.shape[0], D2.data_ptr(), D2.shape[0],
                                           A.data_ptr(), A.shape[1], *inv_params)
        else:
            rc = L.clp_score_pointnormal_dev(self._h, D1.data_ptr(), D1.shape[0], D2.data_ptr(), D2.shape[0],
                                             A.data_ptr(), A.shape[1], *inv_params)
        _capi.check(self._h, rc)
        m = self._m()
        if self._connected_for != m:
            self._connect()
            self._connected_for = m

    def solve(self, u0=None):
        """collective solve; a host barrier first, so that the in-kernel peer waits only ever cover kernel skew"""
        if u0 is None:
            raise ValueError("a sharded solve needs an explicit u0 (every rank must start from the same vector)")
        perm = getattr(self, "_perm", None)
        if perm is not None:
            u0p = np.empty(len(perm), dtype=np.float64); u0p[perm] = np.asarray(u0, dtype=np.float64)
        else:
            u0p = u0
        self._dist.barrier(group=self._group)
        super().solve(u0p)
        if perm is not None:
            s = self._soln
            s.nodes, s.u = unpermute_solution(s.nodes, s.u, perm)
            s.nodes = s.nodes.tolist()
            s.u0 = np.asarray(u0, dtype=np.float64).copy()

    def count_nonzeros(self):
        import torch
        a, b = super().count_nonzeros()
        t = torch.tensor([a, b], dtype=torch.int64,
                         device="cuda" if self._dist.get_backend(self._group) == "nccl" else "cpu")
        self._dist.all_reduce(t, group=self._group)
        return int(t[0]), int(t[1])


class ShardGroup:
    """`world` shards driven from one process, one host thread per shard during solve()."""

    def __init__(self, make_invariant, params, devices, storage=_capi.STORE_F32, same_device=False):
        self.world = len(devices)
        self.shards = []
        L = _capi.load()
        for r, dev in enumerate(devices):
            c = CLIPPER(make_invariant(), params, device=dev, storage=storage)
            _capi.check(c.handle, L.clp_shard_config(c.handle, r, self.world))
            if same_device:
                # the shards' persistent kernels must be co-resident on the one GPU: 1 CTA/SM each for the segmented
                # kernels, and SMs/world fat CTAs each for the resident-vector kernel (it takes a whole SM's shared memory)
                import torch
                sms = torch.cuda.get_device_properties(dev).multi_processor_count
                _capi.check(c.handle, L.clp_set_ctas_per_sm(c.handle, 1))
                _capi.check(c.handle, L.clp_set_grid_cap(c.handle, max(1, sms // self.world)))
            self.shards.append(c)
        self._connected_for = None

    def _connect(self):
        L = _capi.load()
        nb = int(L.clp_shard_blob_bytes())
        blobs = []
        for c in self.shards:
            b = C.create_string_buffer(nb)
            _capi.check(c.handle, L.clp_shard_export(c.handle, b, nb, None))
            blobs.append(b.raw)
        raw = b"".join(blobs)
        for c in self.shards:
            _capi.check(c.handle, L.clp_shard_import(c.handle, raw, nb, self.world))

    def score_pairwise_consistency(self, D1, D2, A=None):
        for c in self.shards:
            c.score_pairwise_consistency(D1, D2, A)
        m = self.shards[0]._m()
        if self._connected_for != m:
            self._connect()
            self._connected_for = m

    def set_matrix_data(self, M, Cm):
        for c in self.shards:
            c.set_matrix_data(M, Cm)
        m = self.shards[0]._m()
        if self._connected_for != m:
            self._connect()
            self._connected_for = m

    def solve(self, u0):
        """collective solve: the persistent kernels of all shards must run concurrently"""
        errs = [None] * self.world

        def work(r):
            try:
                self.shards[r].solve(u0)
            except Exception as e:  # noqa: BLE001
                errs[r] = e

        ts = [threading.Thread(target=work, args=(r,)) for r in range(self.world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        for e in errs:
            if e is not None:
                raise e
        return [c.get_solution() for c in self.shards]

    def count_nonzeros(self):
        tot = [0, 0]
        for c in self.shards:
            a, b = c.count_nonzeros()
            tot[0] += a; tot[1] += b
        return tuple(tot)


def gathered_matvec(partial_rows, group=None):
    """The exchange step written with a collective, for cross-checks: every rank contributes the
    y-entries of its own row block; all_gather (equivalently: all_reduce of zero-padded vectors,
    which is exact because each entry has exactly one non-zero contributor) rebuilds y."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=partial_rows.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([partial_rows.numel()], dtype=torch.int64, device=partial_rows.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    pad = max(sizes)
    mine = torch.zeros(pad, dtype=partial_rows.dtype, device=partial_rows.device)
    mine[: partial_rows.numel()] = partial_rows
    outs = [torch.empty(pad, dtype=partial_rows.dtype, device=partial_rows.device) for _ in range(world)]
    dist.all_gather(outs, mine, group=group)
    return torch.cat([o[:n] for o, n in zip(outs, sizes)])


# ------------------------------------------------------------------------------------------
# bench.py, N > 1
# ------------------------------------------------------------------------------------------
def run_bench(args, METRIC, UNIT):
    """strong scaling of the SAME problem (configs[1], m=20000 by default) over N row shards;
    device time, max over ranks.  The extra key `config4` holds BASELINE.json's m=80000 case."""
    import json
    import os
    import time
    import torch
    import torch.distributed as dist
    from . import datagen
    import clipper_b200 as clipperpy

    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    # torch and the library share one real stream: the legacy default stream's handle is 0, which clp_set_stream reads
    # as "create your own", and the timing events must sit on the stream the kernels run on
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))

    def one_workload(name, m_override, steps, warmup):
        prob = datagen.config_problem(name, m_override)
        cfg = prob["cfg"]; m = cfg["m"]
        # byte-balanced shards: the generator puts the (denser) inlier associations last, so equal ROW counts give the last
        # of 8 shards 1.14x the mean bytes; dealing the associations to the shards in proportion to their sizes (a
        # relabelling of the same problem, see interleave_permutation) brings that to 1.01x.  CLP_SHARD_INTERLEAVE=0 keeps
        # the generator's order.
        if os.environ.get("CLP_SHARD_INTERLEAVE") != "0":
            prob["A"], prob["u0"] = permute_problem(prob["A"], prob["u0"], interleave_permutation(m, world))
        ip = clipperpy.invariants.EuclideanDistanceParams(); ip.sigma, ip.epsilon = cfg["sigma"], cfg["epsilon"]
        clip = ShardedCLIPPER(clipperpy.invariants.EuclideanDistance(ip), clipperpy.Params())
        if os.environ.get("CLP_DENSE_MODE"):
            clip.set_dense_mode(int(os.environ["CLP_DENSE_MODE"]))
        if os.environ.get("CLP_CTAS_PER_SM"):
            _capi.check(clip.handle, _capi.load().clp_set_ctas_per_sm(clip.handle, int(os.environ["CLP_CTAS_PER_SM"])))
        stream = torch.cuda.current_stream()   # run_bench installed a non-default stream (handle != 0)
        clip.set_stream(stream.cuda_stream)
        D1 = torch.from_numpy(np.ascontiguousarray(prob["D1"].T)).to(dev)
        D2 = torch.from_numpy(np.ascontiguousarray(prob["D2"].T)).to(dev)
        A = torch.from_numpy(np.ascontiguousarray(prob["A"].T)).to(dev)
        u0 = torch.from_numpy(prob["u0"]).to(dev)
        u_out = torch.empty_like(u0)
        nodes = np.zeros(m, np.int32)
        sol = _capi.ClpSolution()
        L = _capi.load(); h = clip.handle

        def step():
            clip.score_device("euclidean", D1, D2, A, cfg["sigma"], cfg["epsilon"], 0.0)
            _capi.check(h, L.clp_solve_dev(h, u0.data_ptr(), C.byref(sol), u_out.data_ptr(),
                                           nodes.ctypes.data_as(C.POINTER(C.c_int32))))

        for _ in range(max(warmup, 3)):
            step()
        torch.cuda.synchronize(); dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kms, prof = [], []
        ev0.record(stream)
        for _ in range(steps):
            step()
            kms.append(sol.kernel_ms)
            prof.append((sol.prof_matvec_ms, sol.prof_combine_ms, sol.prof_exchange_ms))
        ev1.record(stream)
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([ev0.elapsed_time(ev1), float(np.mean(kms))], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # every rank must have produced the identical solution
        chk = torch.tensor([sol.score, float(sol.n_nodes), float(sol.n_evals)], dtype=torch.float64, device=dev)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "ranks disagree on the solution"
        # ---- end to end: the same step through the host-pointer C-ABI calls, pinned host buffers in, Solution out,
        #      copies inside the timed region; wall clock, max over ranks.  Guarded: a failure here must not cost
        #      the device-timed line.
        e2e_ms, e2e_note, h2d = None, None, 0
        try:
            if os.environ.get("CLP_BENCH_E2E_MULTI") == "0":
                raise RuntimeError("disabled (CLP_BENCH_E2E_MULTI=0)")
            hD1 = torch.from_numpy(np.ascontiguousarray(prob["D1"].T)).pin_memory()
            hD2 = torch.from_numpy(np.ascontiguousarray(prob["D2"].T)).pin_memory()
            hA = torch.from_numpy(np.ascontiguousarray(prob["A"].T)).pin_memory()
            hu0 = torch.from_numpy(np.ascontiguousarray(prob["u0"])).pin_memory()
            hu = torch.empty(m, dtype=torch.float64).pin_memory()
            dp = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_double))
            ipt = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_int32))
            sol_h = _capi.ClpSolution()
            n1, n2 = prob["D1"].shape[1], prob["D2"].shape[1]

            def step_host():
                _capi.check(h, L.clp_score_euclidean(h, dp(hD1), 3, n1, dp(hD2), n2, ipt(hA), m,
                                                     cfg["sigma"], cfg["epsilon"], 0.0))
                _capi.check(h, L.clp_solve(h, dp(hu0), C.byref(sol_h), dp(hu),
                                           nodes.ctypes.data_as(C.POINTER(C.c_int32)), None))

            for _ in range(2):
                step_host()
            torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                step_host()
            torch.cuda.synchronize(); dist.barrier()
            te = torch.tensor([1e3 * (time.perf_counter() - t0)], dtype=torch.float64, device=dev)
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            e2e_ms = float(te[0])
            h2d = hD1.numel() * 8 + hD2.numel() * 8 + hA.numel() * 4 + m * 8
            if sol_h.score != sol.score or sol_h.n_nodes != sol.n_nodes:
                e2e_note = "host-path solution differs from the device-path one"
        except Exception as e:  # noqa: BLE001
            e2e_ms, e2e_note = None, "host-buffer e2e pass: %s" % (e,)
        ok = torch.tensor([1.0 if e2e_ms is not None else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok[0]) == 0.0:
            e2e_ms = None
        mode = clip.dense_mode()
        if mode in (3, 6):
            kept, pass_bytes = clip.sparse_info()      # this rank's rows
        else:
            r0, nrows = shard_rows(m, rank, world)
            kept, pass_bytes = None, 4 * nrows * m
        pb = torch.tensor([float(pass_bytes)], dtype=torch.float64, device=dev)
        dist.all_reduce(pb, op=dist.ReduceOp.MAX)
        return dict(m=m, ms=float(t[0]), kernel_ms=float(t[1]), n_matvec=int(sol.n_matvec), n_evals=int(sol.n_evals),
                    mode=mode, pass_bytes=float(pb[0]),
                    F=float(sol.score), n_nodes=int(sol.n_nodes), cfg=cfg, e2e_ms=e2e_ms, e2e_note=e2e_note, h2d=h2d,
                    phase_ms=dict(zip(("dense_passes", "combine", "exchange"), np.mean(prof, axis=0).tolist())))

    from bench import ClockSampler, measured_peaks
    sampler = ClockSampler(torch.cuda.current_device()); sampler.start()
    main = one_workload(args.workload, args.m, args.steps, args.warmup)
    clocks = sampler.stop()
    extra = None
    if not getattr(args, "no_config4", False) and args.workload == "c2" and args.m is None:
        extra = one_workload("c4", None, max(1, min(3, args.steps)), 1)
    peak, peak_src = measured_peaks()
    m = main["m"]
    value = m * args.steps / (main["ms"] * 1e-3)
    alg = main["n_matvec"] * main["pass_bytes"]   # algorithmic bytes per GPU per launch (largest shard)
    ach = alg / (main["kernel_ms"] * 1e-3) / 1e9
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": main["ms"] / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 (f32 affinity storage, fp64 vectors/accumulators/decisions)", "data": "synthetic",
            "config": {"workload": "%s: synthetic EuclideanDistance m=%d, 95%% outliers, M row-sharded over %d GPUs, "
                                   "in-kernel NVLink peer-memory exchange" % (args.workload, m, world),
                       "l2": "per-GPU dense slice of M = %.2f GB" % (4.0 * m * m / world / 1e9),
                       "sweep_mode": main["mode"], "algorithmic_bytes_per_pass_per_gpu": main["pass_bytes"],
                       "evals_per_solve": main["n_evals"], "solver_kernel_ms": main["kernel_ms"],
                       "solver_phase_ms": main["phase_ms"], "F": main["F"],
                       "n_nodes": main["n_nodes"]},
            "clocks": clocks,
            "e2e": ({"value": m * args.steps / (main["e2e_ms"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(main["h2d"]),
                     "d2h_bytes_per_step": int(m * 8 + 320), "ms_per_step": main["e2e_ms"] / args.steps,
                     "note": "per rank: D1, D2, A, u0 from pinned host memory, Solution back; wall clock, max over ranks"
                             + ("; " + main["e2e_note"] if main["e2e_note"] else "")}
                    if main["e2e_ms"] else
                    {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(m * 8 + 320),
                     "note": "N>1: inputs replicated in HBM on every rank; result D2H inside the timed region"
                             + ("; " + main["e2e_note"] if main["e2e_note"] else "")}),
            "gpu_launches": (9 if main["mode"] in (3, 6) else 3) * args.steps * world,  # per rank: gather, score, solver (+6 building the compact copy)
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": None, "kernel": "solver_kernel<float>, per GPU", "peak_source": peak_src},
        }
        if extra:
            m4 = extra["m"]
            line["config4"] = {"workload": "c4: m=%d row-sharded over %d GPUs" % (m4, world),
                               "value": m4 * max(1, min(3, args.steps)) / (extra["ms"] * 1e-3), "unit": UNIT,
                               "solver_kernel_ms": extra["kernel_ms"], "evals": extra["n_evals"],
                               "solver_phase_ms": extra["phase_ms"],
                               "per_gpu_gbs": extra["n_matvec"] * extra["pass_bytes"] / (extra["kernel_ms"] * 1e-3) / 1e9,
                               "sweep_mode": extra["mode"],
                               "F": extra["F"], "n_nodes": extra["n_nodes"]}
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
