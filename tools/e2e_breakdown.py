"""Where the host-to-host time of one bench step goes: kernel with / without curve streaming, configure, scalars, plain D2H."""
import sys, time; sys.path.insert(0, '.')
import numpy as np, covid19abm_b200 as cv
from bench import CFG, seeds_for, N_AGENTS, T_DAYS
R = 1000
ip = cv.ModelParameters(**CFG)
pop = cv.Population(R, N_AGENTS, T_DAYS, store_hmatrix=True)
shape = (R, 6, T_DAYS, 12)
pin_inc = cv.PinnedArray(shape, np.int16); pin_prev = cv.PinnedArray(shape, np.int16)
def t(f, n=10):
    f(); t0 = time.perf_counter()
    for _ in range(n): f()
    return 1e3 * (time.perf_counter() - t0) / n
for mode in ("no streaming", "streaming"):
    pop.stream_curves_to(*( (pin_inc.array, pin_prev.array) if mode == "streaming" else (None, None)))
    ms = []
    for k in range(8):
        pop.configure(ip, seeds_for(k, 0, R)); pop.run(sync=True); ms.append(pop.last_run_ms)
    print(mode, "kernel ms", np.round(ms[2:], 3))
    print("  configure ms", round(t(lambda: pop.configure(ip, seeds_for(1, 0, R))), 3))
    def step():
        pop.configure(ip, seeds_for(3, 0, R)); pop.run(sync=False); pop.curves(pin_inc.array, pin_prev.array); pop.scalars()
    print("  full step ms", round(t(step), 3))
    def step2():
        pop.configure(ip, seeds_for(3, 0, R)); pop.run(sync=True)
    print("  configure+run(sync) ms", round(t(step2), 3))
    pop.run(sync=True)
    print("  curves() alone ms", round(t(lambda: pop.curves(pin_inc.array, pin_prev.array)), 3), " scalars() alone ms", round(t(lambda: pop.scalars()), 3))
# main(ip) returns the full state matrix (:141): what its read-back costs for the whole batch (Int16, widened on the device)
hm = cv.PinnedArray((R, T_DAYS, N_AGENTS), np.int16)
pop.run(sync=True)
print("hmatrix_all() ms (1000 x 500 x 10000 Int16 = 10 GB over PCIe)", round(t(lambda: pop.hmatrix_all(hm.array), 2), 1))
hm.free(); hm8 = cv.PinnedArray((R, T_DAYS, N_AGENTS), np.uint8)
print("hmatrix_all(uint8) ms (5 GB)", round(t(lambda: pop.hmatrix_all(hm8.array), 2), 1))
print("hmatrix(0) ms (one replicate, 10 MB)", round(t(lambda: pop.hmatrix(0), 5), 3))
