#!/usr/bin/env python
"""Where the end-to-end step (TopdownPoseEstimator.test_step on pinned host crops) spends its time."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import probpose_code_b200.mmpose_api as api
from probpose_code_b200 import synth
B = 64
model = api.MODELS.build(api.probpose_small_cfg(precision="fp16x3"))
model.load_state_dict(synth.make_state_dict(seed=0)); model.to("cuda:0")
samples = api.make_data_samples(B)
host = [synth.make_crops(B, seed=i).pin_memory() for i in range(4)]
for i in range(3):
    model.test_step(dict(inputs=host[i % 4], data_samples=samples))
torch.cuda.synchronize()
N = 20
t0 = time.perf_counter()
for i in range(N):
    model.test_step(dict(inputs=host[i % 4], data_samples=samples))
t_all = (time.perf_counter() - t0) / N
dev = [h.cuda() for h in host]
eng = model._fused_engine(2 * B, dev[0].device)
rec = torch.empty((B, 17, 7), device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(N):
    eng.infer(dev[i % 4], flip_test=True, flip_indices=api.COCO_FLIP_INDICES, out=rec)
torch.cuda.synchronize(); t_gpu = (time.perf_counter() - t0) / N
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(N):
    x = host[i % 4].to("cuda:0", non_blocking=True); torch.cuda.synchronize()
t_h2d = (time.perf_counter() - t0) / N
t0 = time.perf_counter()
for i in range(N):
    r = rec.cpu()
t_d2h = (time.perf_counter() - t0) / N
t0 = time.perf_counter()
for i in range(N):
    preds = model.head.pack_records(rec)
t_pack = (time.perf_counter() - t0) / N
t0 = time.perf_counter()
for i in range(N):
    model.add_pred_to_datasample(model.head.pack_records(rec), None, samples)
t_pack2 = (time.perf_counter() - t0) / N
t0 = time.perf_counter()
for i in range(N):
    api.PoseDataPreprocessor.stack(host[i % 4])
t_stack = (time.perf_counter() - t0) / N
print(f"test_step {t_all*1e3:.3f} ms | engine.infer (resident) {t_gpu*1e3:.3f} | H2D {t_h2d*1e3:.3f} | D2H rec {t_d2h*1e3:.3f} | pack_records {t_pack*1e3:.3f} | pack+add_pred {t_pack2*1e3:.3f} | stack {t_stack*1e3:.3f}")
