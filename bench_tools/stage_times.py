"""Device time of the two stages of the C2 chain on their own and pipelined (CUDA graphs, CUDA events).
    python bench_tools/stage_times.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def graph_time(fn, reps=50):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = fn()  # noqa: F841
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    dev = torch.device("cuda", 0)
    chain = bench.Chain(dev, seed=1002)
    chain.upload()
    chain.prime()

    def geo():
        cur = torch.cuda.current_stream()
        chain.geo_stream.wait_stream(cur)
        with torch.cuda.stream(chain.geo_stream):
            t = chain.geometry(chain.d["pos"])
        cur.wait_stream(chain.geo_stream)
        return t

    print("geometry stage alone   %8.1f us" % graph_time(geo))
    print("feature stage alone    %8.1f us" % graph_time(lambda: chain.features(chain.tables)))
    print("serial (geo, features) %8.1f us" % graph_time(lambda: chain.step_serial()[0]))
    print("pipelined step         %8.1f us" % graph_time(chain.step))


if __name__ == "__main__":
    main()
