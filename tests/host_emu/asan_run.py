"""Driver of tests/test_host_emu_asan.py: runs the emulated MSM pipeline from a sanitizer-instrumented build
(AddressSanitizer + UBSan) over awkward shapes; compute-sanitizer is closed on the GPU pool, so out-of-bounds
indexing in the pipeline logic is hunted here instead."""
import ctypes, sys, random, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import zkt_oracle as O
from tests import util as U
lib=ctypes.CDLL(sys.argv[1])
u32p=ctypes.POINTER(ctypes.c_uint32); u8p=ctypes.POINTER(ctypes.c_uint8)
ptr=lambda a,t=u32p: a.ctypes.data_as(t) if a is not None else None
rnd=random.Random(4)
dl=[rnd.randrange(1,O.R) for _ in range(40)]
pts=[O.scalar_mul(O.G1_GEN,k) for k in dl]
xy,inf=U.g1_points_to_array(pts)
ok=True
for rounds in ("0","2"):
  os.environ["ZKMSM_BATCH_ROUNDS"]=rounds; os.environ["ZKMSM_BATCH_T"]="3"
  for (c,pre,L,K,n) in [(0,0,0,0,40),(4,0,3,2,40),(6,3,3,2,37),(9,1,0,0,40),(13,0,0,0,5),(5,2,2,4,1),(3,0,1,1,40)]:
    sc=[rnd.randrange(O.R) for _ in range(n)]
    s=U.scalars_to_array(sc); out=np.zeros(24,dtype=np.uint32); oi=ctypes.c_uint32(0)
    rc=lib.emu_g1_msm(ptr(xy),None,ptr(s),n,len(pts),c,pre,L,K,ptr(out),ctypes.byref(oi))
    got=U.g1_from_array(out,oi.value); exp=U.expected_from_dlogs(O.G1_GEN,dl[:n],sc)
    ok = ok and rc==0 and got==exp
# the transform-based quotient (fr_ntt.cuh): product tree, Newton inverse, batched transforms with awkward n
for n in (2, 3, 7, 33, 70):
  u,v,w=([rnd.randrange(O.R) for _ in range(n)] for _ in range(3))
  out=np.zeros((max(n-1,1),8),dtype=np.uint32); tt=np.zeros((n+1,8),dtype=np.uint32); flag=ctypes.c_uint32(0)
  lib.emu_fr_quotient_ntt(ptr(U.scalars_to_array(u)),ptr(U.scalars_to_array(v)),ptr(U.scalars_to_array(w)),n,ptr(out),ctypes.byref(flag),ptr(tt))
  out2=np.zeros((max(n-1,1),8),dtype=np.uint32); flag2=ctypes.c_uint32(0)
  lib.emu_fr_quotient(ptr(U.scalars_to_array(u)),ptr(U.scalars_to_array(v)),ptr(U.scalars_to_array(w)),n,ptr(out2),ctypes.byref(flag2))
  ok = ok and out[:n-1].tolist()==out2[:n-1].tolist() and (flag.value!=0)==(flag2.value!=0)
print("asan run ok:", ok)
