"""Developer probe (ISR_PHASE_LOG build of libisr.so): where a CTA of the verification search spends
its cycles -- Chamfer verification of a few candidates, per-CTA phase stamps.
    ISR_LIBISR_PATH=.../variants/phase.so python scripts/probe_phases_verify.py [candidates]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from imagesequenceregistrationfor6dposeestimationlabeling_b200 import _lib, api, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
lib = _lib.load()
cloud = synth.make_cloud(100000, 1)
R_true, _ = synth.true_pose(3)
Rs, _, k0 = synth.make_candidates(1000, 10, R_true=R_true, t_true=np.zeros(3))
Mq, Mt = synth.verification_matrices(Rs[:B], R_true)
cd, Mqd, Mtd = api._points(cloud, dev), api._poses(Mq, dev), api._poses(Mt, dev)
api.verify_poses(cd, Mqd, Mtd)
torch.cuda.synchronize()
cap = 2 * (2 * B * 400 + 8)
log = torch.zeros((cap, 4), dtype=torch.int64, device="cuda")
lib.isr_debug_cta_log(ctypes.c_void_p(log.data_ptr()), cap)
api.verify_poses(cd, Mqd, Mtd)
torch.cuda.synchronize()
lib.isr_debug_cta_log(None, 0)
L = log.cpu().numpy().astype(np.uint64)
L2 = L[cap // 2:cap - 1]
L = L[:cap // 2 - 1]
keep = L[:, 0] > 0
L, L2 = L[keep], L2[:len(keep)][keep]
lo, sh = np.uint64(0xFFFFFFFF), np.uint64(32)
f = lambda a: a.astype(np.float64)
search_end, q_ready, hints = f(L[:, 0]), f(L[:, 1] & lo), f(L[:, 1] >> sh)
rows, seeds = f(L[:, 2] & lo), f(L[:, 2] >> sh)
total = f(L[:, 3] & np.uint64(0xFFFFF)) * 256
print(f"{len(L)} CTAs logged ({B} candidates x 2 directions x 391 blocks; the last launch of the call)")
def st(name, v):
    print(f"  {name}: mean {v.mean():.0f} ({100 * v.mean() / total.mean():.1f} %)  median {np.median(v):.0f}  p90 {np.percentile(v, 90):.0f}")
st("query copy", q_ready)
st("(hints)", hints - q_ready)
st("row spheres", rows - hints)
st("seed search + queueing", seeds - rows)
st("main loop", search_end - seeds)
st("  exact tests + copy issue", f(L2[:, 0] & lo))
st("  wait for the sub-tile", f(L2[:, 0] >> sh))
st("  filter scan", f(L2[:, 1] & lo))
st("  resolve", f(L2[:, 1] >> sh))
st("  produce (sphere walk -> FIFO)", f(L2[:, 2] & lo))
st("  nearest-first sort", f(L2[:, 2] >> sh))
st("result write-back", total - search_end)
st("whole CTA", total)
nsc, nte = f(L2[:, 3] >> np.uint64(48)), f((L2[:, 3] >> sh) & np.uint64(0xFFFF))
npa, nqu = f((L2[:, 3] >> np.uint64(16)) & np.uint64(0xFFFF)), f(L2[:, 3] & np.uint64(0xFFFF))
print(f"  per CTA: {nsc.mean():.1f} scanned tiles, {nqu.mean():.1f} quarter scans, {nte.mean():.1f} exact tests, {npa.mean():.1f} resolve passes")
print(f"  per event (mean cycles): exact test {f(L2[:, 0] & lo).sum() / nte.sum():.0f}, quarter scan {f(L2[:, 1] & lo).sum() / nqu.sum():.0f}, "
      f"resolve pass {f(L2[:, 1] >> sh).sum() / npa.sum():.0f}")
