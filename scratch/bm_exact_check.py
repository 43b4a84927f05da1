"""Is the device Box-Muller bit-identical to oracle/rng.py box_muller?  (needs a GPU)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genjax_b200.runtime import smc_ops
from oracle import rng
dev = torch.device("cuda", 0)
words = (0x12345678, 0x9ABCDEF0)
n = 1 << 20
z = smc_ops.normal_fill(words, 0, 3, n, 4, dev).cpu().numpy()
o = rng.normal_vec(words, np.arange(n, dtype=np.uint64), 3, 4)
bad = np.argwhere(z.view(np.uint32) != o.view(np.uint32))
print("mismatching normals:", len(bad), "of", z.size, "max abs diff", np.abs(z - o).max())
w = np.stack(rng.site_words(words, np.arange(n, dtype=np.uint64), 3, 0), 1)
for i, j in bad[:8]:
    b0, b1 = w[i, 2 * (j // 2)], w[i, 2 * (j // 2) + 1]
    print(i, j, hex(b0), hex(b1), "u1", rng.u01(np.uint32(b0)), "u2", rng.u01(np.uint32(b1)), "dev", z[i, j], z[i, j].view(np.uint32), "oracle", o[i, j], o[i, j].view(np.uint32))
