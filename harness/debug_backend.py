import sys, os
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests')]
import numpy as np
import orc
from harness import synth, pipeline
from vins_mono_b200 import Estimator

seq = synth.Sequence(seed=11, duration=6.0)
NMSG = int(sys.argv[1]) if len(sys.argv) > 1 else 20
kw = dict(window_size=20, estimate_td=1, tr=0.033) if len(sys.argv) > 2 and sys.argv[2] == 'w20' else {}
max_feats = 300 if kw else 150
msgs = synth.track_messages(seq, NMSG, max_feats=max_feats)
cpu, gpu = orc.OracleEstimator(orc.be_config(**kw)), Estimator(tic=synth.TIC, ric=synth.RIC, **kw)
t_imu, acc, gyr = seq.imu()
fa, fb = pipeline.ImuFeeder(t_imu, acc, gyr), pipeline.ImuFeeder(t_imu, acc, gyr)
seeds = pipeline.gt_seed_rows(seq, [m[0] for m in msgs])
cpu.set_seed(seeds, seq.ba, seq.bg); gpu.set_seed(seeds, seq.ba, seq.bg)
for stamp, ids, d in msgs:
    fa.feed(cpu, stamp); fb.feed(gpu, stamp)
    cpu.processImage(ids, d, stamp); gpu.processImage(ids, d, stamp)
    ia, ib = cpu.info(), gpu.info()
    if ia['solver_flag'] != 1: continue
    sa, _ = cpu.states(); sb, _ = gpu.states()
    Aa, ba_, blka = cpu.prior(); Ab, bb_, blkb = gpu.prior()
    perm = []
    for (t, i, off, sz) in blkb:
        o = next(x for x in blka if x[0] == t and (t >= 2 or x[1] == i))
        perm += list(range(o[2], o[2] + sz))
    Aa, ba_ = Aa[np.ix_(perm, perm)], ba_[perm]
    sc = np.sqrt(np.outer(np.abs(np.diag(Aa)) + 1e-12, np.abs(np.diag(Aa)) + 1e-12))
    print(f"t={stamp:.1f} marg={ia['marginalization_flag']}/{ib['marginalization_flag']} it={ia['iterations']}/{ib['iterations']} ok={ia['successful_steps']}/{ib['successful_steps']} term={ia['termination']}/{ib['termination']} "
          f"cost0 {ia['initial_cost']:.6f}/{ib['initial_cost']:.6f} cost {ia['final_cost']:.6f}/{ib['final_cost']:.6f} dp={np.abs(sa[:,0:3]-sb[:,0:3]).max():.2e} dv={np.abs(sa[:,7:10]-sb[:,7:10]).max():.2e} "
          f"n={len(ba_)}/{len(bb_)} dA={np.abs(Aa-Ab).max()/np.abs(Aa).max():.2e} dArel={(np.abs(Aa-Ab)/sc).max():.2e} db={np.abs(ba_-bb_).max()/np.abs(ba_).max():.2e} t={gpu.timing()} dbg={gpu.solver_debug()}")
