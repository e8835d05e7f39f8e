"""Why does the loopback rate per GPU drop when two instances run side by side?  On a 2-GPU box:
  A  both generators at once, one shared MPS daemon (what bench.py does)
  B  GPU 0 alone under the shared daemon
  C  both at once, one MPS daemon per GPU (each daemon and its clients see only their GPU)
  D  both at once, shared daemon, no CPU pinning
"""
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def run_pair(boxes, devs, label, seconds=1.0):
    out = {}

    def one(box, dev, arg):
        g = box.generator(arg, 200, 10, min_seconds=seconds)
        out[dev] = round(g["infer_per_s"] / 1e3, 1)

    ths = [threading.Thread(target=one, args=(boxes[d], d, a)) for d, a in devs]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    print(label, out, "sum", round(sum(out.values()), 1), flush=True)


class PerGpuBox(bench.LoopbackBox):
    """One daemon per GPU: the daemon, the server and the generator of GPU d run with CUDA_VISIBLE_DEVICES=d."""

    def __init__(self, dev):
        super().__init__([0], use_mps=True, manage_mps=True)
        self.phys = dev
        tag = "_g%d" % dev
        self.MPS_ENV = {k: v + tag for k, v in bench.LoopbackBox.MPS_ENV.items()}
        self.env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(dev), **self.MPS_ENV)

    def start_mps(self):
        for d in self.MPS_ENV.values():
            os.makedirs(d, exist_ok=True)
        subprocess.run(["nvidia-cuda-mps-control", "-d"], env=self.env, timeout=30, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        time.sleep(1.0)

    def stop_mps(self):
        subprocess.run(["nvidia-cuda-mps-control"], input="quit\n", env=self.env, text=True, timeout=30, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def main():
    with bench.LoopbackBox([0, 1], use_mps=True, manage_mps=True) as box:
        run_pair({0: box, 1: box}, [(0, 0), (1, 1)], "warm", 0.3)
        run_pair({0: box, 1: box}, [(0, 0), (1, 1)], "A shared daemon, both  ")
        run_pair({0: box}, [(0, 0)], "B shared daemon, GPU0  ")
        run_pair({1: box}, [(1, 1)], "B shared daemon, GPU1  ")
        run_pair({0: box, 1: box}, [(0, 0), (1, 1)], "A again                ")
    time.sleep(2.0)
    b0, b1 = PerGpuBox(0), PerGpuBox(1)
    with b0, b1:
        run_pair({0: b0, 1: b1}, [(0, 0), (1, 0)], "warm", 0.3)
        run_pair({0: b0, 1: b1}, [(0, 0), (1, 0)], "C daemon per GPU, both ")
        run_pair({0: b0}, [(0, 0)], "C daemon per GPU, GPU0 ")
    time.sleep(2.0)
    with bench.LoopbackBox([0, 1], use_mps=True, manage_mps=True, pin=False) as box:
        run_pair({0: box, 1: box}, [(0, 0), (1, 1)], "warm", 0.3)
        run_pair({0: box, 1: box}, [(0, 0), (1, 1)], "D shared, unpinned     ")


if __name__ == "__main__":
    main()
