"""Event trace of one persistent attention CTA (debug build: EXTRA=-DSELFTOK_ATTN_TRACE profiles/mk_variant.sh trace .../attn_tc5.cu):

    SELFTOK_B200_LIB=build/ab/lib_trace.so python profiles/attn_trace.py [S] > gpurun_out/attn_trace.json

tags: 1 softmax warp passed s_full | 2 S loaded from TMEM | 3 exponentials + row max done | 4 P stored (tcgen05.wait::st) |
      10 MMA warp passed p_ready | 11 P V issued + committed | 12 Q K^T issued + committed.  Clock = SM cycles (low 32 bits).
"""
import ctypes as C
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from selftoktokenizer_b200 import capi  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 768
GEN = int(sys.argv[2]) if len(sys.argv) > 2 else 5          # 5: attn_tc5 (ns code 10), 6: attn_tc6 (ns code 20)
dev = torch.device("cuda:0")
lib = capi.load_library()
dump = lib.selftok_dbg_attn_trace6 if GEN == 6 else lib.selftok_dbg_attn_trace
dump.argtypes = [C.c_void_p, C.c_int]
NS = 0          # IEEE-half single pass
B, H = 16, 24
qkv = torch.randn(B, S, 3, H, 64, device=dev)
buf = (C.c_ulonglong * 65536)()
capi.k_attention_tc(qkv, H, NS)
dump(buf, 65536)          # discard the warm-up launch
capi.k_attention_tc(qkv, H, NS)
n = dump(buf, 65536)
ev = [(int(buf[i] >> 56), int((buf[i] >> 48) & 0xff), int((buf[i] >> 32) & 0xffff), int(buf[i] & 0xffffffff)) for i in range(n) if buf[i]]
print(json.dumps({"S": S, "B": B, "H": H, "events": ev}))
