#!/usr/bin/env python
"""Times the decode kernel (planted / model-like logits, batch 256) - used with PP_DECODE_DBG stage cut-offs."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cases
from probpose_code_b200 import ops
from tools.kernel_bench import time_ms
fi = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for name, gen in (("planted", lambda s: cases.planted_peak_logits(batch, seed=s)), ("noise1", lambda s: cases.noise_logits(batch, s, 1.0))):
    z = torch.from_numpy(gen(1)).cuda(); zf = torch.from_numpy(gen(2)).cuda()
    m0, b0 = time_ms(lambda: ops.decode(z, input_is_logits=True))
    m1, b1 = time_ms(lambda: ops.decode(z, zf, fi, input_is_logits=True))
    print(json.dumps(dict(dbg=os.environ.get("PP_DECODE_DBG", "0"), inputs=name, batch=batch, plain_us=m0 * 1e3, tta_us=m1 * 1e3)))
m, b = time_ms(lambda: torch.amax(z, dim=-1))
print(json.dumps(dict(ref="torch.amax over the same 53.6 MB", us=m * 1e3)))
