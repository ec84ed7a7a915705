"""CPU baseline workers for bench.py: the reference's viterbi_mem + traceback_viterbi (oracle/_ref, compiled
unmodified with -O3 -DNDEBUG; the C port if it is absent), ONE PROCESS PER CORE, each on its
own contiguous block of the seeded pair stream (pairs are i.i.d. draws, so a block keeps the bin weights).
Workers are separate interpreters (`python tools/cpu_worker.py ...`): they import neither torch nor the product."""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def work(workload, seed, first, n, table, g, e, k, start_at, lib=None):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import ctypes as C
    import oracle
    from synth import synth_pairs
    table = np.ascontiguousarray(table, dtype=np.float32)
    w = synth_pairs(n, workload, seed, first, threads=1)
    a_off, b_off = w["a_off"], w["b_off"]
    cells = float((np.diff(a_off).astype(np.float64) * np.diff(b_off).astype(np.float64)).sum())
    total = int(a_off[-1] + b_off[-1]) + n
    out_a, out_b = np.zeros(total + 1, np.uint8), np.zeros(total + 1, np.uint8)
    out_len, score = np.zeros(n, np.uint64), np.zeros(n, np.float32)
    vp = lambda x: C.c_void_p(x.ctypes.data)  # noqa: E731
    ref = C.CDLL(lib) if lib else oracle.ref
    while time.time() < start_at:     # common start: the slowest worker's time is the wall time of the pool
        time.sleep(0.001)
    if ref is not None:
        fn = ref.coati_ref_viterbi_batch
        fn.restype = C.c_double
        secs = fn(C.c_size_t(n), vp(w["a_all"]), vp(a_off), vp(w["b_all"]), vp(b_off), vp(w["anc_all"]), vp(w["des_all"]),
                  vp(table), C.c_float(g), C.c_float(e), C.c_size_t(k), C.c_int(1), vp(out_a), vp(out_b), vp(out_len),
                  vp(score))
        kind = "reference"
    else:
        t0 = time.perf_counter()
        for p in range(n):
            sa = slice(int(a_off[p]), int(a_off[p + 1]))
            sb = slice(int(b_off[p]), int(b_off[p + 1]))
            oracle.viterbi(w["anc_all"][sa].tobytes().decode(), w["des_all"][sb].tobytes().decode(), table, g, e, k,
                           enc=(w["a_all"][sa], w["b_all"][sb]))
        secs = time.perf_counter() - t0
        kind = "port"
    return dict(cells=cells, seconds=float(secs), pairs=n, kind=kind)


def run_pool(workload, seed, first, n, table, g, e, k, procs, lib=None):
    """n pairs from `first`, dealt to `procs` single-threaded processes in contiguous blocks.  `lib`: another build
    of the reference shim (oracle/_ref/libcoati_ref_o2g.so: Meson's default -O2 -g with assertions)."""
    procs = max(1, min(procs, n))
    bounds = [first + n * i // procs for i in range(procs + 1)]
    with tempfile.TemporaryDirectory() as tmp:
        tpath = os.path.join(tmp, "table.npy")
        np.save(tpath, np.ascontiguousarray(table, dtype=np.float32))
        start_at = time.time() + 3.0 + 0.05 * procs      # after every process has imported and generated its block
        env = dict(os.environ, OMP_NUM_THREADS="1")
        ps = []
        for i in range(procs):
            spec = dict(workload=workload, seed=seed, first=bounds[i], n=bounds[i + 1] - bounds[i], table=tpath,
                        g=g, e=e, k=k, start_at=start_at, lib=lib)
            ps.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), json.dumps(spec)],
                                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
        res = []
        for p in ps:
            out, err = p.communicate(timeout=900)
            if p.returncode != 0:
                raise RuntimeError("CPU worker failed: " + err[-500:])
            res.append(json.loads(out.strip().splitlines()[-1]))
    if any(r["seconds"] <= 0 for r in res):
        raise RuntimeError("CPU reference run failed")
    return res


if __name__ == "__main__":
    s = json.loads(sys.argv[1])
    print(json.dumps(work(s["workload"], s["seed"], s["first"], s["n"], np.load(s["table"]), s["g"], s["e"], s["k"],
                          s["start_at"], s.get("lib"))))
