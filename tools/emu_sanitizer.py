#!/usr/bin/env python
"""The device code of the stepping kernels, compiled for the host (tests/emu) with gcc's
-fsanitize=address,undefined, over the cases that exercise its indexing: the unrolled kernel's
sector buffer with blocks of rows that straddle sectors, the kernel for one delay equation
with rings small enough to wrap.  compute-sanitizer is closed on this GPU pool
(profiles/r2_memcheck_refused.txt); this and the index-checking debug build
(tools/lag_stats.py) stand in.  usage:
  (cd tests/emu && DDE_EMU_EXTRA="-fsanitize=address,undefined -fno-omit-frame-pointer" \
      DDE_EMU_OUT=libdde_emu_asan.so bash build.sh)
  LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tools/emu_sanitizer.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from dde_b200 import workloads
from emu import pyemu
from oracle import pyoracle
def same(a,b): return np.array_equal(a.view(np.uint64), b.view(np.uint64))
for n_times, ri in [(12,False),(12,True),(14,True),(102,True),(2,False),(2,True),(1001,False)]:
    model,y0,parms,times,kw=workloads.lorenz_sweep(17,n_times=n_times,t_end=2.0)
    kw=dict(kw,return_initial=ri)
    got=pyemu.dopri_batch(model,y0,times,parms,mode=1,tag="_asan",**kw)
    want=pyoracle.dopri_batch("oracle",model,y0,times,parms,**kw)
    assert same(got.y,want.y) and np.array_equal(got.statistics,want.statistics), (n_times,ri)
model,y0,parms,times,kw=workloads.mackey_glass(96)
got=pyemu.dopri_batch(model,y0,times,parms,mode=4,tag="_asan",**kw)
want=pyoracle.dopri_batch("oracle",model,y0,times,parms,**kw)
assert same(got.y,want.y)
for nh in (3,8):
    kw2=dict(kw,n_history=nh)
    got=pyemu.dopri_batch(model,y0[:32],times,parms[:32],mode=4,tag="_asan",**kw2)
    want=pyoracle.dopri_batch("oracle",model,y0[:32],times,parms[:32],**kw2)
    assert same(got.y,want.y) and np.array_equal(got.code,want.code)
print("asan/ubsan run clean: unrolled kernel's sector rows, dde1 kernel incl. small rings")
