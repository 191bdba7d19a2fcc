"""Small finite-difference applies in CUDA-graph replay over rotating buffers: streaming kernels vs the gather kernel."""
import os
import sys
import statistics

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numgrids_b200 as ng  # noqa: E402

A = ng.AxisType
for shape in ((512, 512), (1024, 1024), (2048, 2048), (128, 128, 128), (256, 256, 256)):
    g = ng.Grid(*(ng.create_axis(A.EQUIDISTANT, n, 0.0, 1.0) for n in shape))
    pts = int(np.prod(shape))
    nbuf = max(4, min(80, (320 << 20) // (16 * pts)))
    ins = torch.randn(nbuf, *shape, dtype=torch.float64, device="cuda")
    outs = torch.empty_like(ins)
    for axis in range(len(shape)):
        op = ng.Diff(g, 1, axis).operator.operator
        res = []
        for mode, env in (("stream", {"NG_FD_STREAM": "1", "NG_FD_STREAM_MIN": "0"}), ("gather", {"NG_FD_STREAM": "0"})):
            os.environ.update(env)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(s):
                for i in range(nbuf):
                    op.apply_ptr(ins[i].data_ptr(), outs[i].data_ptr(), stream=s.cuda_stream)
                torch.cuda.synchronize()
                with torch.cuda.graph(graph, stream=s):
                    for rep in range(3):
                        for i in range(nbuf):
                            op.apply_ptr(ins[i].data_ptr(), outs[i].data_ptr(), stream=s.cuda_stream)
            torch.cuda.current_stream().wait_stream(s)
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); graph.replay(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) / (3 * nbuf))
            res.append(f"{mode} {statistics.median(ts) * 1e3:7.2f} us")
        print(f"{shape} axis {axis}: " + "   ".join(res), flush=True)
    del ins, outs
