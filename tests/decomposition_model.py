"""Work-decomposition model of the blend kernels on the CPU oracle's tile lists -- TEST INFRASTRUCTURE (imports oracle/).

Counts, for a sample of tiles of a workload, how many warp iterations each pixel-to-lane decomposition of the forward /
backward blend needs (the round-1 kernels were instruction-issue bound, so iterations x instructions per iteration is the
cost model; VERDICT r01 item 2 asked for the decomposition to be changed, not the knobs):

  A        round-1 forward: 4x2-pixel blocks, four blocks in lock step per warp, per 32-entry window
  A_free   same blocks, free-running through a 256-entry batch
  B        one pixel per lane, lock step per 32-entry window            <- blend_fwd.cu (round 2)
  C        one pixel per lane, free-running through a batch (C512 / C1024 / C1048576: larger batches)
  bwd_q    round-1 backward: per-block hit words, free-running through a batch
  bwd_lane one pixel per lane over EXACT per-pixel hit bits, free-running through a batch   <- blend_bwd.cu phase 1

usage: python tests/decomposition_model.py [workload] [tiles]      (prints iterations per warp; ~1 min at headline_1m)
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsorb_slam_b200.scene import make_config
from oracle import gs_oracle
name = sys.argv[1] if len(sys.argv) > 1 else 'headline_1m'
ntiles = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sc = make_config(name)
t0 = time.time()
fr = gs_oracle.frame_from_scene(sc)
print('oracle fwd', time.time() - t0, 's; R', fr.num_rendered)
g = fr.geometry(); b = fr.binning(); im = fr.image_state()
W, H = sc.cam.width, sc.cam.height
tx, ty = (W + 15) // 16, (H + 15) // 16
m2 = g['means2D']; co = g['conic_opacity']
# cull data as preprocess.cu computes it (need cov_xx, cov_yy: invert conic)
A, B, Cc, op = co[:, 0].astype(np.float64), co[:, 1].astype(np.float64), co[:, 2].astype(np.float64), co[:, 3].astype(np.float64)
det = A * Cc - B * B
with np.errstate(all='ignore'):
    cov_xx = Cc / det; cov_yy = A / det
    lim = np.log(255.0 * op)
    t2 = 2 * (lim * 1.002 + 0.01)
    thr = -0.5 * t2
    ex = np.sqrt(t2 * cov_xx) * 1.02 + 0.05
    ey = np.sqrt(t2 * cov_yy) * 1.02 + 0.05
rng = np.random.default_rng(1)
tiles = rng.choice(tx * ty, ntiles, replace=False)
tot = dict(A=0, B=0, C=0, Cq=0, useful=0, cand=0, entries=0, batches=0, warps=0, A_free=0, cullwin=0, bwd_lane=0, bwd_useful=0, bwd_q=0)
for t in tiles:
    s, e = b['ranges'][t]
    ids = b['point_list'][s:e]
    n = len(ids)
    if n == 0: continue
    X0, Y0 = (t % tx) * 16, (t // tx) * 16
    pxs = (X0 + np.arange(16))[None, :].repeat(16, 0).reshape(-1).astype(np.float32)
    pys = (Y0 + np.arange(16))[:, None].repeat(16, 1).reshape(-1).astype(np.float32)
    inside = (pxs < W) & (pys < H)
    x = m2[ids, 0][:, None]; y = m2[ids, 1][:, None]
    dx = x - pxs[None]; dy = y - pys[None]
    a_, b_, c_, o_ = co[ids, 0][:, None], co[ids, 1][:, None], co[ids, 2][:, None], co[ids, 3][:, None]
    power = -0.5 * (a_ * dx * dx + c_ * dy * dy) - b_ * dx * dy
    alpha = np.minimum(0.99, o_ * np.exp(power))
    valid = (power <= 0) & (alpha >= 1 / 255.0)
    Tafter = np.cumprod(np.where(valid, 1 - alpha, 1.0), axis=0)
    stop = valid & (Tafter < 1e-4)
    done_idx = np.where(stop.any(0), stop.argmax(0), n)   # entry index at which the pixel becomes done
    done_idx = np.where(inside, done_idx, -1)
    idx = np.arange(n)[:, None]
    alive = idx <= done_idx[None]          # pixel still walks entry idx (the done-triggering entry is evaluated)
    blended = valid & (idx < done_idx[None])
    cand_px = (np.abs(dx) <= ex[ids][:, None]) & (np.abs(dy) <= ey[ids][:, None]) & alive
    assert (blended & ~cand_px).sum() == 0
    # tile walk length: until all pixels done
    walk = min(n, int(done_idx.max()) + 1)
    tot['entries'] += walk
    tot['useful'] += int(blended.sum()); tot['cand'] += int(cand_px.sum())
    nb = (walk + 255) // 256
    tot['batches'] += nb
    # pixel -> warp region (8x4), lane
    lx = (np.arange(256) % 16); ly = (np.arange(256) // 16)
    warp = (ly // 4) * 2 + (lx // 8)
    quarter = ((ly % 4) // 2) * 2 + ((lx % 8) // 4)
    for w in range(8):
        pm = warp == w
        wd = int(done_idx[pm].max()) + 1      # warp walks entries < wd
        if wd <= 0: continue
        wd = min(wd, n)
        tot['warps'] += 1
        cw = cand_px[:wd][:, pm]                 # [entries, 32]
        bw = blended[:wd][:, pm]
        qidx = quarter[pm]
        # block-level candidates (current scheme): block bbox test
        bx0 = X0 + (w & 1) * 8; by0 = Y0 + (w >> 1) * 4
        nwin = (wd + 31) // 32
        tot['cullwin'] += nwin
        pad = nwin * 32 - wd
        # per quarter block masks
        qc = np.zeros((nwin * 32, 4), bool)
        for q in range(4):
            xa0 = bx0 + (q & 1) * 4; xa1 = xa0 + 3; ya0 = by0 + (q >> 1) * 2; ya1 = ya0 + 1
            xs = m2[ids[:wd], 0]; ys = m2[ids[:wd], 1]; exx = ex[ids[:wd]]; eyy = ey[ids[:wd]]
            hit = ~((xs + exx < xa0) | (xs - exx > xa1)) & ~((ys + eyy < ya0) | (ys - eyy > ya1))
            qdone = done_idx[pm][qidx == q].max()
            hit &= (np.arange(wd) <= qdone)
            qc[:wd, q] = hit
        qcw = qc.reshape(nwin, 32, 4).sum(1)          # [win, 4]
        tot['A'] += int(qcw.max(1).sum())
        # free running quarters per batch
        nbw = (wd + 255) // 256
        qcb = np.zeros((nbw * 8, 4), int); qcb[:nwin] = qcw
        tot['A_free'] += int(qcb.reshape(nbw, 8, 4).sum(1).max(1).sum())
        cwp = np.zeros((nwin * 32, 32), bool); cwp[:wd] = cw
        lw = cwp.reshape(nwin, 32, 32).sum(1)        # [win, lane]
        tot['B'] += int(lw.max(1).sum())
        # B with exact masks (only the pairs that pass the power test reach the walk), and B over 64-entry double windows
        bwp0 = np.zeros((nwin * 32, 32), bool); bwp0[:wd] = (valid & alive)[:wd][:, pm]
        tot['B_exact'] = tot.get('B_exact', 0) + int(bwp0.reshape(nwin, 32, 32).sum(1).max(1).sum())
        n2 = (nwin + 1) // 2
        z = np.zeros((n2 * 64, 32), bool); z[:wd] = cw
        tot['B64'] = tot.get('B64', 0) + int(z.reshape(n2, 64, 32).sum(1).max(1).sum())
        z = np.zeros((n2 * 64, 32), bool); z[:wd] = (valid & alive)[:wd][:, pm]
        tot['B64_exact'] = tot.get('B64_exact', 0) + int(z.reshape(n2, 64, 32).sum(1).max(1).sum())
        lb = np.zeros((nbw * 8, 32), int); lb[:nwin] = lw
        tot['C'] += int(lb.reshape(nbw, 8, 32).sum(1).max(1).sum())
        for BS in (512, 1024, 1 << 20):
            k = 'C%d' % BS
            nbb = (wd + BS - 1) // BS
            if BS >= 1 << 20:
                v = int(cw.sum(0).max())
            else:
                z = np.zeros((nbb * BS, 32), bool); z[:wd] = cw
                v = int(z.reshape(nbb, BS, 32).sum(1).max(1).sum())
            tot[k] = tot.get(k, 0) + v
            k = 'bl%d' % BS
            if BS >= 1 << 20:
                v = int(bw.sum(0).max())
            else:
                z = np.zeros((nbb * BS, 32), bool); z[:wd] = bw
                v = int(z.reshape(nbb, BS, 32).sum(1).max(1).sum())
            tot[k] = tot.get(k, 0) + v
        # backward: per lane walk of blended bits per batch (free running), vs quarter visits (hit words)
        bwp = np.zeros((nwin * 32, 32), bool); bwp[:wd] = bw
        blb = np.zeros((nbw * 8 * 32, 32), bool); blb[:nwin * 32] = bwp
        tot['bwd_lane'] += int(blb.reshape(nbw, 256, 32).sum(1).max(1).sum())
        qh = np.stack([bwp[:, qidx == q].any(1) for q in range(4)], 1)   # [entries,4]
        qhb = np.zeros((nbw * 256, 4), bool); qhb[:nwin * 32] = qh
        tot['bwd_q'] += int(qhb.reshape(nbw, 256, 4).sum(1).max(1).sum())
        tot['bwd_useful'] += int(bw.sum())
        # other lane-group shapes for the backward: (bw x bh)-pixel blocks, one block per group of bw*bh lanes, free-running per batch
        lxw = lx[pm] % 8; lyw = ly[pm] % 4
        for bw_, bh_ in ((2, 2), (4, 1), (2, 1), (8, 1), (4, 4), (8, 2)):
            gid = (lyw // bh_) * (8 // bw_) + (lxw // bw_)
            ng = (8 // bw_) * (4 // bh_)
            gh = np.stack([bwp[:, gid == g_].any(1) for g_ in range(ng)], 1)
            ghb = np.zeros((nbw * 256, ng), bool); ghb[:nwin * 32] = gh
            k = 'bwd_%dx%d' % (bw_, bh_)
            tot[k] = tot.get(k, 0) + int(ghb.reshape(nbw, 256, ng).sum(1).max(1).sum())
            tot[k + '_visits'] = tot.get(k + '_visits', 0) + int(gh.sum())
        tot['bwd_q_mean'] = tot.get('bwd_q_mean', 0) + qh.sum() / 4.0          # perfectly balanced quarters
        tot['bwd_q_tile'] = tot.get('bwd_q_tile', 0) + int(qh.sum(0).max())   # free-running through the whole tile list
nt = len(tiles)
w = tot['warps']
print({k: v / nt for k, v in tot.items()})
print({k: v / w for k, v in tot.items() if k[0] in 'Cb'})
print('per warp: A(cur fwd lockstep)=%.1f A_free=%.1f B(lane,window)=%.1f C(lane,batch)=%.1f cullwin=%.1f | bwd quarter=%.1f bwd lane=%.1f useful/lane-slot fwd C=%.2f bwd lane=%.2f' % (
    tot['A'] / w, tot['A_free'] / w, tot['B'] / w, tot['C'] / w, tot['cullwin'] / w, tot['bwd_q'] / w, tot['bwd_lane'] / w,
    tot['useful'] / (tot['C'] * 32), tot['bwd_useful'] / (tot['bwd_lane'] * 32)))
