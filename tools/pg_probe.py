"""Where does the pose-graph solve spend its time?  C5 graph, several back-to-back solves, SM clock sampled by NVML."""
import sys, time, threading
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from rgbdslam_v2_b200 import Frontend, synth
from rgbdslam_v2_b200._capi import default_params
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
clk = []
stop = threading.Event()
def poll():
    while not stop.is_set():
        clk.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
        stop.wait(0.005)
th = threading.Thread(target=poll, daemon=True); th.start()
p = default_params(); p.depth_cov_z0 = 2.0
fe = Frontend(0, p)
for nv, ne in ((5000, 30000), (2000, 21000)):
    g = synth.make_pose_graph(nv, ne, seed=0)
    time.sleep(1.0)  # idle GPU
    for it in range(4):
        t0 = time.perf_counter()
        x, chi2, lm, cg = fe.optimize_graph(g["init"], g["fixed"], g["ij"], g["meas"], g["info"], stop=0.01)
        t1 = time.perf_counter()
        cs = [c for t, c in clk if t0 <= t <= t1]
        print(f"nv {nv} call {it}: {1e3*(t1-t0):.1f} ms, lm {lm}, pcg {cg}, us/pcg-iter {1e6*(t1-t0)/max(cg,1):.1f}, SM MHz min/med/max {min(cs)}/{sorted(cs)[len(cs)//2]}/{max(cs)}")
stop.set()
