"""Small-n batches through the general schedule: evaluations/s and n^3-flop rate."""
import sys, time, numpy as np, torch
sys.path[:0] = [".", "oracle", "tests"]
import helpers
import waveome_b200 as wb
from waveome_b200.engine import Engine, Batch
eng = Engine(0)
st = torch.cuda.ExternalStream(eng.stream)
for n, B in [(60, 20000), (100, 20000), (127, 20000), (150, 10000), (250, 5000)]:
    X, y = helpers.make_data(n, seed=n)
    rng = np.random.default_rng(0)
    Y = y[None, :] + 0.1 * rng.normal(size=(B, n))
    m = wb.GPR(helpers.saturated_kernel(), mean_function=wb.ConstantMean())
    bt = Batch(eng, X, Y, [m.program()])
    x = bt.x0()
    bt.eval(x)
    bt.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    xd = torch.tensor(x, device="cuda"); fd = torch.empty(B, dtype=torch.float64, device="cuda"); gd = torch.empty_like(xd)
    ld = torch.empty_like(fd); sd = torch.empty(B, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(3): bt.eval_device(xd, fd, gd, ld, sd)
        e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    prof = {k: round(v[0] / 3, 3) for k, v in bt.profile_read().items() if v[0] > 0}
    print("n=%d B=%d: %.3f ms/eval-batch -> %.0f evals/s, n^3 rate %.2f TFLOP/s (%.1f%% of 35.5); classes %s" % (
        n, B, ms, B / ms * 1e3, B * n**3 / ms / 1e9, 100 * B * n**3 / ms / 1e9 / 35.5, prof), flush=True)
    bt.close()
