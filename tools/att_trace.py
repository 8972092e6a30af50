"""Clock stamps of the warp-specialised attention kernel (CTA 0, softmax group 0, first 4 units): per tile
[wait S begin, S ready, softmax done, O ready, read-out done] in cycles relative to the first stamp.
    python tools/att_trace.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    trace = torch.zeros(128, dtype=torch.int64, device="cuda")
    os.environ["PP_ATT_TRACE"] = hex(trace.data_ptr())
    from probpose_code_b200 import _lib, ops
    B, N, H, DH = 128, 192, 12, 32
    prec = _lib.PRECISIONS["fp16x3"]
    torch.manual_seed(0)
    bufs = [ops.to_operand(torch.randn(B * N, 3 * H * DH, device="cuda"), prec) for _ in range(3)]
    o = ops.attention(bufs[0], B, N, H, DH, prec, impl=2)
    for i in range(4):
        ops.attention(bufs[i % 3], B, N, H, DH, prec, impl=2, out=o)
    torch.cuda.synchronize()
    tm = trace.cpu()[64:96].view(8, 4)
    t = trace.cpu()[:64].view(8, 8)[:, :5]
    t0 = int(t[0, 0])
    names = ["wait_S", "softmax", "wait_O", "readout"]
    for r in range(8):
        row = [int(x) - t0 for x in t[r]]
        print(f"unit {r // 2} tile {r % 2}: start {row[0]:7d} | " + " ".join(f"{n} {row[k + 1] - row[k]:6d}" for k, n in enumerate(names))
              + f" | MMA thread: chunk0 seen {int(tm[r, 0]) - t0:7d}, last chunk seen {int(tm[r, 1]) - t0:7d}, issued {int(tm[r, 2]) - t0:7d}; softmax done at {row[2]:7d}, O ready at {row[3]:7d}")


if __name__ == "__main__":
    main()
