import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_classification_icbhi_b200 import LogMelPlan
plan = LogMelPlan(device="cuda:0"); dev = plan.device
n = 6900
rs = np.random.RandomState(0)
lens = (np.clip(rs.lognormal(np.log(2.5), 0.5, n), 0.2, 16.2) * 16000).astype(np.int64)
starts = np.concatenate([[0], np.cumsum((lens + 3) // 4 * 4)[:-1]])
wave = torch.randn(int(starts[-1] + lens[-1]) + 4, device=dev) * 0.1
off = torch.from_numpy(starts).to(dev); ln = torch.from_numpy(lens.astype(np.int32)).to(dev)
ref = plan.forward(wave, off, ln).clone()
ok = True
for i in range(30):
    plan.set("max_ctas", int(rs.choice([148, 148, 100, 37, 7])))
    o = plan.forward(wave, off, ln)
    torch.cuda.synchronize()
    ok &= bool(torch.equal(o, ref))
print("ragged corpus, 30 launches with varying grid sizes, bit-identical:", ok)
