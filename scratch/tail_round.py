"""Where does a tail round (few active models) spend its time: host enqueue, device chain, or the per-round sync?"""
import sys, time, numpy as np, torch
sys.path[:0] = [".", "oracle", "tests"]
import helpers
import waveome_b200 as wb
from waveome_b200.engine import Engine, Batch
eng = Engine(0)
st = torch.cuda.ExternalStream(eng.stream)
n = 500
X, y = helpers.make_data(n, seed=5)
rng = np.random.default_rng(0)
for B in (1, 8, 64):
    Y = y[None, :] + 0.3 * rng.normal(size=(B, n))
    k = wb.SquaredExponential(active_dims=[1]) + wb.Categorical(active_dims=[0]) * wb.SquaredExponential(active_dims=[2])
    m = wb.GPR(k, mean_function=wb.ConstantMean())
    bt = Batch(eng, X, Y, [m.program()])
    x = bt.x0()
    xd = torch.tensor(x, device="cuda"); fd = torch.empty(B, dtype=torch.float64, device="cuda"); gd = torch.empty_like(xd)
    ld = torch.empty_like(fd); sd = torch.empty(B, dtype=torch.int32, device="cuda")
    for _ in range(5): bt.eval_device(xd, fd, gd, ld, sd)
    torch.cuda.synchronize()
    N = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(st)
    for _ in range(N): bt.eval_device(xd, fd, gd, ld, sd)
    e1.record(st)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("B=%d: enqueue %.1f us/eval (host), device %.1f us/eval, wall incl. drain %.1f us/eval" % (
        B, (t1 - t0) / N * 1e6, e0.elapsed_time(e1) / N * 1e3, (t2 - t0) / N * 1e6), flush=True)
    # synchronised evaluations (what a fit round does: enqueue, then wait)
    t0 = time.perf_counter()
    for _ in range(N):
        bt.eval_device(xd, fd, gd, ld, sd); torch.cuda.synchronize()
    print("      synchronised: %.1f us/eval" % ((time.perf_counter() - t0) / N * 1e6), flush=True)
    bt.profile(True)
    t0 = time.perf_counter(); r = bt.fit(); dt = time.perf_counter() - t0
    c = bt.counters()
    prof = {k_: (round(v[0], 1), int(v[1])) for k_, v in bt.profile_read().items() if v[0] > 0}
    print("      fit (profiling on): %.3f s, %d rounds -> %.1f us/round; classes (ms, launches): %s" % (dt, c["rounds"], dt / max(1, c["rounds"]) * 1e6, prof))
    bt.profile(False)
    t0 = time.perf_counter(); r = bt.fit(); dt = time.perf_counter() - t0
    c2 = bt.counters()
    print("      fit: %.3f s, %d rounds -> %.1f us/round" % (dt, c2["rounds"] - c["rounds"], dt / max(1, c2["rounds"] - c["rounds"]) * 1e6))
    bt.close()
