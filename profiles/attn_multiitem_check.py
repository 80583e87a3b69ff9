"""Multi-item correctness sweep of the tcgen05 attention kernels (persistent CTAs walking several items each):
    python profiles/attn_multiitem_check.py [ns_code] S1 S2 ...      (ns 10: attn_tc5, 20: attn_tc6)"""
import os
import sys
import torch
import torch.nn.functional as F
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from selftoktokenizer_b200 import capi  # noqa: E402
ns = int(sys.argv[1])
dev = torch.device("cuda:0")
B, H = int(os.environ.get("ATTN_CHECK_B", "16")), 24
for S in [int(x) for x in sys.argv[2:]]:
    g = torch.Generator().manual_seed(S)
    qkv = torch.randn(B, S, 3, H, 64, generator=g).to(dev)
    try:
        o = capi.k_attention_tc(qkv, H, ns)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"S={S}: FAILED {str(e)[:200]}", flush=True)
        break
    q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, S, H * 64)
    print(f"S={S}: max-abs err {float((o - ref).abs().max()):.3e}", flush=True)
