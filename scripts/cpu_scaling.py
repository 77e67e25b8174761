import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench
from problems import make_batch_fast
import multiprocessing as mp
ss, way, vlim, alim = make_batch_fast(2048, seed=1)
grid = np.linspace(0, 1, 200)
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for n in (1, 8, 32, 64, 128):
    pool = mp.get_context("fork").Pool(n, initializer=bench._ref_worker_init)
    S = min(2048, 64 * n)
    idx = np.array_split(np.arange(S), n * 4)
    tasks = [(ss, way[i], vlim[i], alim[i], grid) for i in idx if len(i)]
    pool.map(bench._ref_solve_chunk, tasks[:n], chunksize=1)
    t0 = time.perf_counter(); pool.map(bench._ref_solve_chunk, tasks, chunksize=1); dt = time.perf_counter() - t0
    print("procs %3d: %6.0f paths/s  (%.1f per proc)" % (n, S / dt, S / dt / n))
    pool.close(); pool.join()
