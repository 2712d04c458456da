"""Developer check: workload-generator data (C3 / C4 shapes at a chosen scale) through the emulated kernels (tests/native/emu) against the oracle.
   python tools/emu_workload_check.py <subscriptions> <topics> [prebuilt libemu.so]"""
import sys, time, ctypes as C
ROOT = __import__('pathlib').Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import numpy as np
import test_emu_kernels as T
from oracle import oracle as orc
from rmqtt_b200 import workload as wl
lib = C.CDLL(sys.argv[3]) if len(sys.argv) > 3 else C.CDLL(str(T._build(__import__('pathlib').Path(__import__('tempfile').mkdtemp()), san=False)))
lib.emu_new.restype = C.c_void_p
for f in ("emu_sub_add", "emu_sub_remove", "emu_retain_set", "emu_retain_remove", "emu_match", "emu_retain_match"):
    getattr(lib, f).restype = C.c_int32
ns, nt = int(sys.argv[1]), int(sys.argv[2])
cfg = wl.C3.scaled(n_subs=ns, n_topics=nt)
sb, so, sv = wl.gen_subs(cfg)
tb, to = wl.gen_topics(cfg)
e, tree = T.Emu(lib), orc.TopicTree()
t0=time.time()
blob=sb.tobytes()
for i in range(ns):
    f=blob[so[i]:so[i+1]]
    e.lib.emu_sub_add(e.h, f, len(f), int(sv[i]), 0)
tree.bulk_insert(sb, so, sv)
print('build', time.time()-t0)
for flags in (4, 1):
    t0=time.time()
    res, work, deferred = e.match(tb, to, flags)
    want = tree.match_batch(tb, to)
    T._same(res, want)
    c = want['counters']
    print('flags', flags, 'ok', time.time()-t0, 'deferred', deferred, [int(x) for x in work], [c['V'],c['E'],c['F'],c['M']])
rcfg = wl.C4.scaled(n_subs=ns, n_topics=max(200, nt//10))
rb, ro, rv = wl.gen_retained(rcfg); fb, fo = wl.gen_retain_filters(rcfg)
rt = orc.RetainTree(); rt.bulk_insert(rb, ro, rv)
rblob=rb.tobytes()
for i in range(len(rv)):
    t=rblob[ro[i]:ro[i+1]]
    e.lib.emu_retain_set(e.h, t, len(t), int(rv[i]))
t0=time.time()
res, work, grew = e.retain_match(fb, fo, stats=1)
T._same(res, rt.match_batch(fb, fo))
print('retained ok', time.time()-t0, 'grew', grew, [int(x) for x in work], 'hits', res.needed)
