"""Many small problems side by side.

The reference's own use case is m <= 2000 associations per registration (benchmarks/main.cpp:206,254-270,
README.md:85): one such problem cannot fill a B200, and the persistent solver then spends its time in device-wide
barriers between hundreds of CTAs (config 1, m = 1000: 46 evaluations of 22 us each).  The reference's threading
contract already says how to batch: distinct clipper::CLIPPER objects are independent.  So a batch is
a pool of CLIPPER objects, each with

  * its own CUDA stream (every handle owns one) and its own device workspace,
  * a cap on the CTAs of its persistent kernels (clp_set_grid_cap), so that the cooperative launches of different
    objects are resident at the same time on disjoint SMs and each synchronises a handful of CTAs only,
  * one host thread driving it (the C-ABI calls block; ctypes releases the GIL while they run).

EXPERIMENTAL: written at the end of round 1 after the GPU budget was spent -- the device code it drives is the
validated solver, but grids this small and concurrent cooperative launches have not been run on hardware yet.
tests/test_gpu_batch.py is therefore opt-in (CLP_TEST_EXPERIMENTAL=1)."""
from __future__ import annotations

import queue
import threading

from . import api


class BatchSolver:
    """pool of `workers` CLIPPER objects sharing one GPU; `grid_cap` CTAs each (0 = no cap)"""

    def __init__(self, invariant_factory, params=None, workers=16, grid_cap=24, device=0, dense_mode=None):
        self._objs = []
        for _ in range(int(workers)):
            c = api.CLIPPER(invariant_factory(), params if params is not None else api.Params(), device=device)
            if dense_mode is not None:
                c.set_dense_mode(dense_mode)
            c.set_grid_cap(grid_cap)
            self._objs.append(c)

    @property
    def workers(self):
        return len(self._objs)

    def solve_many(self, problems):
        """problems: iterable of dicts with D1, D2, A (or None) and optionally u0.
        Returns the list of api.Solution in input order (plus .associations: the selected (k, 2) pairs)."""
        problems = list(problems)
        todo = queue.SimpleQueue()
        for k, p in enumerate(problems):
            todo.put((k, p))
        out = [None] * len(problems)
        errors = []

        def run(clip):
            while True:
                try:
                    k, p = todo.get_nowait()
                except queue.Empty:
                    return
                try:
                    clip.score_pairwise_consistency(p["D1"], p["D2"], p.get("A"))
                    clip.solve(p.get("u0"))
                    s = clip.get_solution()
                    s.associations = clip.get_selected_associations()
                    out[k] = s
                except Exception as e:  # keep the other workers going; re-raised below
                    errors.append((k, e))

        threads = [threading.Thread(target=run, args=(c,), daemon=True) for c in self._objs]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            k, e = errors[0]
            raise RuntimeError("problem %d of the batch failed: %s" % (k, e)) from e
        return out
