"""How long do the NVML queries of bench.py's ClockSampler take, and what do they cost a launch-latency-bound loop?"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
x = torch.zeros(1024, device="cuda")
for name, f in (("GetClockInfo", lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                ("GetCurrentClocksEventReasons", lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))):
    ts = []
    for _ in range(50):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    print(f"{name}: median {1e3 * np.median(ts):.3f} ms, max {1e3 * max(ts):.3f} ms")
def loop(n=20000):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        x.add_(1.0)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6
print(f"tiny-kernel launch loop alone: {loop():.2f} us per launch")
for interval in (0.02, 0.1):
    stop = threading.Event()
    def poll():
        while not stop.is_set():
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            time.sleep(interval)
    t = threading.Thread(target=poll, daemon=True); t.start()
    print(f"with the sampler thread polling every {interval * 1e3:.0f} ms: {loop():.2f} us per launch")
    stop.set(); t.join()
