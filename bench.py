#!/usr/bin/env python
"""bench.py -- train drug-pairs/sec of the Bi-GNN step on B200 (+ segment-SpMM HBM roofline).

  python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
  python bench.py --impl reference --steps K --warmup W    # CPU port of the reference path

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'train drug-pairs/sec'
UNIT = 'pairs/s'
LOWER_ONLY_POS = 32768          # positive pairs per step of the lower-level-only workload (scaled pair batch)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='ddi_scaled',
                    help='ddi_scaled = BASELINE config 4 (the configuration the 1/2/4/8-GPU metric is quoted on); '
                         'drugcombo_shape = config 2; drugbank_shape = config 1')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--no-graph', action='store_true', help='eager launches instead of CUDA-graph replay')
    ap.add_argument('--no-l2-flush', action='store_true')
    ap.add_argument('--cpu-sample-steps', type=int, default=6)
    ap.add_argument('--dist-profile', action='store_true',
                    help='N > 1: after the timed region, two eager steps with CUDA events around every C-ABI call and '
                         'every collective (rank 0 reports; collective times include waiting for the slowest rank)')
    ap.add_argument('--skip-cpu', action='store_true', help='no cpu_baseline leg (non-default workloads)')
    ap.add_argument('--skip-gpu-eager', action='store_true', help='no gpu_eager_baseline leg')
    ap.add_argument('--ref-device', default='cpu', help='--impl reference: cpu (the reference arm) or cuda (eager baseline)')
    ap.add_argument('--skip-rooflines', action='store_true', help='no kernel roofline legs (non-default workloads)')
    ap.add_argument('--replicated-upper', action='store_true',
                    help='N > 1: evaluate the upper level on every rank instead of row-partitioning it (GCN only)')
    ap.add_argument('--spmm-rows', type=int, default=8_000_000, help='rows of the >L2 segment-SpMM roofline case')
    return ap.parse_args()


def workload_config(name):
    from bignn_b200 import synthetic as S
    w = S.WORKLOADS[name]
    if name.startswith('mol_1m'):
        # BASELINE config 3: the lower-level-only model (src/train.py:99-107) over large pair batches
        return dict(workload='LL-GNN, lower level only (GIN x5, multi-scale mean readout, MLP scorer 640-80-10-1, BCE, Adam) '
                             'on a synthetic dataset of {} shape'.format(name),
                    name=name, drugs=w['N'], ddi_edges=w['M'], mean_atoms=w['mean_atoms'],
                    node_feat=int(sum(w['groups'])), pos_pairs_per_step=LOWER_ONLY_POS, neg_pairs_per_step=LOWER_ONLY_POS,
                    batch_norm_batch='the merged pair batch (one BatchNorm batch per step, as in the reference)')
    arch = ('GIN x5 lower, multi-scale mean readout, MetaLayer x3 upper (one GAT per interaction edge type, summed), '
            'MLP scorer 128-16-3, CE, Adam') if 'drugcombo' in name else \
        ('GIN x5 lower, mean readout (64-dim upper input), GCN x3 upper, MLP scorer, BCE, Adam' if 'ddi_scaled' in name
         else 'GIN x5 lower, multi-scale mean readout, GCN x3 upper, MLP scorer, BCE, Adam')
    return dict(workload='Bi-GNN ({}) on a synthetic dataset of {} shape'.format(arch, name),
                name=name, drugs=w['N'], ddi_edges=w['M'], mean_atoms=w['mean_atoms'],
                node_feat=int(sum(w['groups'])), pos_pairs_per_step=64, neg_pairs_per_step=64,
                lower_chunk_graphs=128)


def make_workload(name, seed):
    from bignn_b200 import synthetic as S
    # one process per node draws the 20 M-edge graph; the other ranks wait for its cache file
    return S.cached_workload(name, seed, writer=int(os.environ.get('LOCAL_RANK', 0)) == 0)


def workload_flags(name, device='cuda:0'):
    """drugcombo_shape runs the DrugCombo architecture the reference ships for that dataset
    (src/config.py:74-88,120-123,224): one GAT per interaction edge type through MetaLayer, 3-class CE."""
    import bignn_b200 as B
    if name.startswith('mol_1m'):
        return B.make_flags(model='lower_level_gnn', device=device)
    if 'drugcombo' in name:
        return B.make_flags(dataset='drugcombo', higher_level_gnn_type='gat', device=device)
    if 'ddi_scaled' in name:          # BASELINE config 4: 64-dim interaction-graph stage (mean readout, not multi-scale)
        return B.make_flags(node_aggr='avg_pool', device=device)
    return B.make_flags(device=device)


def layer_specs(name):
    f = workload_flags(name)
    return [getattr(f, 'layer_%d' % i) for i in range(1, f.layer_num + 1)]


# --------------------------------------------------------------------------- reference arm
# the reference's per-graph host conversion and Python edge set make the 200 k-drug / 20 M-edge configuration
# infeasible as a bounded sample on the CPU (~30 s and tens of GB of autograd state per
# step): each reference-arm step is ONE FULL TRAIN STEP OF THE SAME WORKLOAD AT 1/10 SIZE, and the reported value is
# the linear extrapolation to full size (the step cost scales with drugs and edges, not with the 128 scored pairs).
REFERENCE_SAMPLE = {'ddi_scaled': ('ddi_scaled_small', 10.0), 'ppi_50k': ('ppi_50k_small', 10.0)}
# lower-level-only model: the step cost is proportional to the pairs scored, so the reference arm runs the same model on
# the same kind of data with a smaller pair batch and its pairs/s compare directly (no extrapolation)
LOWER_ONLY_REFERENCE = ('mol_1m_small', 1024)


def run_reference(args, sample_steps=None, device='cpu'):
    """The reference path on the host cores: oracle/bignn_oracle.py (CPU port; the reference is
    pure Python over un-vendored wheels and cannot travel to the GPU box).  device='cuda' runs the same eager
    code on the GPU (what the reference would do on a B200 today)."""
    import torch
    from oracle import bignn_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    lower_only = args.workload.startswith('mol_1m')
    name, scale = REFERENCE_SAMPLE.get(args.workload, (args.workload, 1.0))
    if lower_only:
        name, scale = LOWER_ONLY_REFERENCE[0], 1.0
    w = make_workload(name, args.seed)
    ds = O.PackedDataset(w)
    specs = O.parse_specs(layer_specs(args.workload))
    state = O.init_params(specs, ds.num_node_feat, num_labels=int(w['num_labels']), seed=8,
                          num_edge_types=len(ds.etypes) + 1)
    if lower_only:
        tr = O.OracleLowerOnlyTrainer(ds, specs, state, LOWER_ONLY_REFERENCE[1], device=device)
    else:
        tr = O.OracleTrainer(ds, specs, state, 64, device=device)
    np.random.seed(8)
    torch.manual_seed(8)
    steps = args.steps if sample_steps is None else sample_steps
    warm = min(args.warmup, 2) if sample_steps is not None else args.warmup
    sync = torch.cuda.synchronize if device != 'cpu' else (lambda: None)
    for _ in range(warm):
        tr.step()
    sync()
    t0 = time.perf_counter()
    pairs = 0
    for _ in range(steps):
        _, p = tr.step()
        pairs += p
    sync()
    dt = time.perf_counter() - t0
    measured = pairs / dt
    what = 'torch {} eager fp32'.format('CUDA' if device != 'cpu' else 'CPU')
    if lower_only:
        sample = ('{} full train steps of the same model on {} ({} drugs) with {} pairs/step instead of {} (the step cost is '
                  'proportional to the pairs scored: pairs/s compare directly), {}, {} host threads').format(
            steps, name, ds.N, pairs // max(steps, 1), 2 * LOWER_ONLY_POS, what, cores)
    elif scale == 1.0:
        sample = '{} full train steps of the same workload ({} pairs/step), {}, {} host threads'.format(
            steps, pairs // max(steps, 1), what, cores)
    else:
        sample = ('{} full train steps of the workload at 1/{:g} size ({}: {} drugs, {} DDI edges; {} pairs/step), {}, '
                  '{} host threads: measured {:.1f} pairs/s = {:.1f} ms/step; value = linear extrapolation to full size '
                  '(measured / {:g}: the step cost scales with drugs and edges, not with the pairs scored)').format(
            steps, scale, name, ds.N, len(ds.ddi_row) // 2, pairs // max(steps, 1), what, cores, measured,
            1e3 * dt / steps, scale)
    return dict(value=measured / scale, ms_per_step=1e3 * dt / steps * scale, cores=cores, steps=steps, sample=sample,
                measured_sample_value=measured, extrapolation_factor=scale)


# --------------------------------------------------------------------------- clocks
class ClockSampler(object):
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, universal_newlines=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i] == 'Active' for r in self.rows)]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


# --------------------------------------------------------------------------- our arm
def spmm_roofline(torch, B, rows, peaks, device):
    """Segment SpMM (GIN aggregation, D=64) on a molecule-like merged graph larger than L2:
    achieved = algorithmic bytes / CUDA-event time per launch."""
    from bignn_b200 import ops, synthetic as S
    G = rows // 30
    atom_ptr, nbr_ptr, nbr_idx, _ = S.molecule_graphs(G, 30.0, seed=3)
    A = int(atom_ptr[-1])
    col = nbr_idx.astype(np.int64) + np.repeat(atom_ptr[:-1].astype(np.int64), np.diff(atom_ptr))[
        np.repeat(np.arange(A), np.diff(nbr_ptr))]
    csr = ops.CSR(torch.as_tensor(nbr_ptr).to(device), torch.as_tensor(col.astype(np.int32)).to(device), A)
    D = 64
    x = torch.randn(A, D, device=device)
    y = torch.empty_like(x)
    for _ in range(3):
        ops.spmm(csr, x, ops.SPMM_GIN, 1.0, out=y)
    reps = 10
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        ops.spmm(csr, x, ops.SPMM_GIN, 1.0, out=y)
        b.record()
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    nnz = csr.nnz
    alg = 4.0 * D * A * 2 + 4.0 * nnz + 4.0 * (A + 1)
    ach = alg / (ms * 1e-3) / 1e9
    peak = peaks.get('hbm_gbs', 6650.0)
    # DRAM bytes per launch of this exact case from the committed `ncu --set full` capture
    # (profiles/r1_summary.md: dram__bytes_read.sum + dram__bytes_write.sum = 2.148 + 2.007 GB)
    traffic = 4.155e9 if abs(A - 8000311) < 1000 else None
    return dict(kernel='k_spmm_v4<16,1,GIN> (bignn_spmm_f32)', bound='hbm', achieved=ach, peak=peak, unit='GB/s',
                frac=ach / peak, traffic=traffic, rows=A, nnz=nnz, D=D, ms_per_launch=ms,
                algorithmic_bytes=alg, peak_source='MEASURED_PEAKS.json' if 'hbm_gbs' in peaks else 'fallback 6650')


def parallelism(eng, world):
    if world == 1:
        return 'single'
    sh = eng.chunk_shards if len(eng.chunk_shards) <= 8 and eng.n_chunks_total < 100 else \
        '{} chunks'.format(eng.n_chunks_total)
    if eng.upper is not None:
        return ('drug-sharded lower level x{} (chunks {}), pooled-row exchange, interaction-graph rows partitioned by '
                'source drug (bounds {}), per-layer all-gather + BatchNorm-statistics all-reduce, replicated scorer'
                .format(world, sh, eng.upper.part.bounds.tolist()))
    return 'drug-sharded lower level x{} (chunks {}), pooled-row all-reduce, replicated upper level'.format(world, sh)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import bignn_b200 as B
    from bignn_b200.engine import BiGNNEngine

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = 'cuda:%d' % local
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
            os.environ['NCCL_DEBUG'] = 'WARN'
        # NCCL prints its version banner on stdout when the first communicator comes up (at WARN level too):
        # point fd 1 at stderr until then, so that stdout carries nothing but the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device(dev))
            t = torch.zeros(1, device=dev)
            dist.all_reduce(t)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    B._lib.load()
    peaks = {}
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        peaks = json.load(open(pk))

    B.set_flags(workload_flags(args.workload, dev))
    w = make_workload(args.workload, args.seed)
    data = B.BiGNNData.from_npz(w, device=dev)
    torch.manual_seed(8)
    np.random.seed(8)            # every rank stages the same pair batches (the upper level is replicated)
    model = B.Model(data).to(dev)
    model.train()
    eng = BiGNNEngine(data, model, use_cuda_graph=not args.no_graph, rank=rank, world=world,
                      partition_upper=False if args.replicated_upper else None)
    sampler = B.RandomSampler(data, 64)

    flush = None if args.no_l2_flush else torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up through the public API (also captures the CUDA graph)
    for _ in range(max(args.warmup, 3)):
        st = eng.train_step(sampler)
    eng.read_loss(st)
    l0 = B._lib.launch_count()

    clocks = ClockSampler(local)
    clocks.start()

    # ---- (1) device-resident: inputs already staged in HBM, K replays, events per step
    st, P = eng.stage_pairs(eng.last_batch.batch_gids, [p.true_label for p in eng.last_batch.pair_list])
    eng.step_staged(st, P)
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    graph, sb, _ = eng._graphs[P] if eng.use_cuda_graph else (None, eng.last_static_batch, None)
    for a, b in evs:
        if flush is not None:
            flush.fill_(1.0)
        a.record()
        if graph is not None:
            graph.replay()
        else:
            eng._device_step(sb)
        b.record()
    barrier()
    dev_ms = float(sum(a.elapsed_time(b) for a, b in evs))
    pairs_dev = P * args.steps

    # ---- (2) end to end through the public API: host sampling, pinned H2D, step, loss D2H
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pairs_e2e = 0
    slots = []
    for _ in range(args.steps):
        if flush is not None:
            flush.fill_(1.0)
        st = eng.train_step(sampler)
        pairs_e2e += eng.last_batch.batch_gids.shape[0]
        slots.append(st)
        if len(slots) >= 3:
            eng.read_loss(slots.pop(0))
    losses = [eng.read_loss(s) for s in slots]
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clk = clocks.stop()

    # ---- max over ranks, whole-job aggregates
    # the job scores P pairs per step whatever the rank count (drugs are sharded, pairs are not)
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0]), float(t[1])
    launches_per_step = getattr(eng, 'launches_per_step', None)

    dist_prof = None
    if world > 1 and args.dist_profile:
        from bignn_b200 import dist as bdist
        eng.last_static_batch = sb
        bdist.TIMING = {}
        prof = kernel_profile(torch, B, eng, steps=2, on_start=lambda: bdist.TIMING.clear())
        coll = bdist.timing_summary(2)
        bdist.TIMING = None
        dist_prof = dict(kernels=prof, collectives={k: dict(ms_per_step=round(v[0], 4), calls_per_step=v[1])
                                                    for k, v in sorted(coll.items(), key=lambda kv: -kv[1][0])})
    if rank != 0:
        return None
    out = dict(metric=METRIC, value=pairs_dev / (dev_ms * 1e-3), unit=UNIT, n_gpus=world, steps=args.steps,
               warmup=max(args.warmup, 3), ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling='strong',
               vs_baseline=None, dtype='f32', data='synthetic', impl='ours',
               config=workload_config(args.workload),
               run=dict(l2='flushed between timed steps (256 MiB write)' if flush is not None
                        else 'not flushed (working set < L2)', cuda_graph=bool(eng.use_cuda_graph),
                        parallelism=parallelism(eng, world), lower_path=getattr(eng, 'lower_path', 'layers')),
               e2e=dict(value=pairs_e2e / (e2e_ms * 1e-3), unit=UNIT, h2d_bytes_per_step=eng.h2d_bytes_per_step,
                        d2h_bytes_per_step=eng.d2h_bytes_per_step, ms_per_step=e2e_ms / args.steps,
                        wall_ms_per_step=wall_ms / args.steps),
               clocks=clk, last_loss=losses[-1] if losses else None)
    if dist_prof is not None:
        out['dist_profile'] = dist_prof
    return out, (torch, B, peaks, dev, eng, data, model)


def run_lower_only(args):
    """BASELINE config 3: the lower-level-only model over large pair batches on one GPU (engine_lower.LowerOnlyEngine)."""
    import torch
    import bignn_b200 as B
    from bignn_b200.engine_lower import LowerOnlyEngine, FastPairSampler
    if int(os.environ.get('WORLD_SIZE', 1)) > 1:
        if int(os.environ.get('RANK', 0)) == 0:
            print(json.dumps(dict(impl='ours', unavailable='the lower-level-only workload runs on one GPU in this round')))
        return None
    dev = 'cuda:0'
    torch.cuda.set_device(0)
    B._lib.load()
    peaks = {}
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    B.set_flags(workload_flags(args.workload, dev))
    data = B.BiGNNData.from_npz(make_workload(args.workload, args.seed), device=dev)
    torch.manual_seed(8)
    np.random.seed(8)
    model = B.Model(data).to(dev)
    model.train()
    eng = LowerOnlyEngine(data, model)
    sampler = FastPairSampler(data, LOWER_ONLY_POS, seed=8)
    flush = None if args.no_l2_flush else torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    from bignn_b200.engine_lower import PairPrefetcher
    # the public API for large pair batches: host threads prepare the coming batches (sampling, negatives, labels, unique
    # drugs) while the device runs; every one of them is prepared INSIDE the timed region of the end-to-end leg
    pre = PairPrefetcher(eng, sampler, seed=8)
    for _ in range(max(args.warmup, 3)):
        loss = eng.train_step_prefetched(pre)
    torch.cuda.synchronize()
    pre.close()                     # (nothing prepared before the end-to-end timer starts is used inside it)
    clocks = ClockSampler(0)
    clocks.start()
    # (1) device-resident: the staged batch (unique drug rows, pair positions, labels) is on the host side of nothing --
    # merged-graph construction from the packed dataset in HBM, forward, backward, Adam
    staged = eng.last_static_batch
    P = staged[1].shape[0]
    n0 = B._lib.launch_count()
    eng._device_step(staged)
    eng.launches_per_step = B._lib.launch_count() - n0
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in evs:
        if flush is not None:
            flush.fill_(1.0)
        a.record()
        eng._device_step(staged)
        b.record()
    torch.cuda.synchronize()
    dev_ms = float(sum(a.elapsed_time(b) for a, b in evs))
    # (2) end to end: positive sampling (DataLoader), vectorised negatives, staging, step, loss read-back
    t0 = time.perf_counter()
    pre = PairPrefetcher(eng, sampler, seed=9)      # a fresh stream: its first batch is prepared inside the timed region
    pairs = 0
    for _ in range(args.steps):
        if flush is not None:
            flush.fill_(1.0)
        loss = float(eng.train_step_prefetched(pre))
        pairs += eng.last_pairs
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    host_workers = pre.workers
    pre.close()
    clk = clocks.stop()
    m = eng.last['merged']
    out = dict(metric=METRIC, value=P * args.steps / (dev_ms * 1e-3), unit=UNIT, n_gpus=1, steps=args.steps,
               warmup=max(args.warmup, 3), ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling='strong',
               vs_baseline=None, dtype='f32', data='synthetic', impl='ours', config=workload_config(args.workload),
               run=dict(l2='flushed between timed steps (256 MiB write)' if flush is not None else 'not flushed',
                        cuda_graph=False, parallelism='single', lower_path=eng.lower_path,
                        merged_graph=dict(graphs=m.G, atoms=m.A, directed_bonds=m.E),
                        samplers='vectorised positives (engine_lower.FastPairSampler) and negatives (fast_negative_pairs): the '
                                 "reference's rules and distributions, not its random stream",
                        host='engine_lower.PairPrefetcher: {} host threads prepare the coming batches while the device '
                             'runs (a batch costs ~160 ms of host time, 7x the device step)'.format(host_workers)),
               e2e=dict(value=pairs / (e2e_ms * 1e-3), unit=UNIT, h2d_bytes_per_step=eng.h2d_bytes_per_step,
                        d2h_bytes_per_step=4, ms_per_step=e2e_ms / args.steps),
               clocks=clk, last_loss=loss)
    return out, (torch, B, peaks, dev, eng, data, model)


def _numel(t):
    return int(t.numel()) if t is not None and hasattr(t, 'numel') else 0


def algorithmic_bytes(name, a):
    """SURVEY 8(d) compulsory traffic of one C-ABI call (every input read once, every output written once; fp32
    features, int32 indices), from the call's own arguments.  None = no convention stated for this entry."""
    try:
        if name == 'bignn_spmm_f32':
            n, d, mode = a[6], a[7], a[8]
            return 8.0 * d * n + 4.0 * _numel(a[1]) + 4.0 * (n + 1) + (4.0 * n if mode == 2 else 0.0)
        if name == 'bignn_spmm_rows_f32':
            n, d = a[6], a[8]
            return 4.0 * d * (n + a[2].shape[0]) + 4.0 * _numel(a[1]) + 4.0 * (n + 1) + 4.0 * a[2].shape[0]
        if name == 'bignn_spmm_planned_rows_f32':
            n, d = a[13], a[15]
            return 4.0 * d * (n + a[9].shape[0]) + 4.0 * _numel(a[1]) + 4.0 * (n + 1) + 4.0 * a[9].shape[0]
        if name == 'bignn_gemm_tc_f32':
            return 4.0 * (a[0] * a[2] + a[0] * a[1] + a[1] * a[2])
        if name == 'bignn_gemm_tc_masked_f32':
            return 4.0 * (a[0] * a[2] + 2.0 * a[0] * a[1] + a[1] * a[2])
        if name == 'bignn_gemm_f32':
            M, N, K = a[2], a[3], a[4]
            return 4.0 * (M * K + M * N + K * N)
        if name == 'bignn_dw_tc_f32':
            return 4.0 * a[0] * (a[1] + a[2])
        if name == 'bignn_bn_seg_fwd':
            return 12.0 * a[6] * a[0].shape[0]
        if name == 'bignn_bn_seg_bwd':
            return 20.0 * a[8] * a[0].shape[0]
        if name == 'bignn_readout_fwd':
            return 4.0 * a[4] * (a[0].shape[0] + a[3])
        if name == 'bignn_gin_layer_fwd':
            # read X once (neighbour rows of a molecule are re-read from L1/L2), write Y (+ Z and T when they are kept
            # for the backward), neighbour ids, row pointers
            rows, din, dout, nnz = a[0], a[1], a[2], a[5]
            din_pad = (din + 3) // 4 * 4
            kept = (din_pad if a[22] is not None else 0) + (dout if a[24] is not None else 0)
            return 4.0 * rows * (din_pad + dout + kept) + 4.0 * nnz + 4.0 * (rows + 1)
        if name == 'bignn_merge_build':
            F, G, A, E = a[4], a[6], a[15], a[16]
            return 4.0 * (A + 1) + 8.0 * E + 4.0 * A + 4.0 * (G + 1) + 8.0 * F * A
    except Exception:
        return None
    return None


def kernel_profile(torch, B, eng, steps=3, on_start=None):
    """Per-entry-point device time of one EAGER step (CUDA events around every C-ABI call on the launching stream;
    a pass of its own after the timed region, not the CUDA-graph replay that `value` times) -- finds the dominant
    kernel of the step, its average launch duration and its algorithmic bytes."""
    lib = B._lib
    rec = []
    orig = lib.call

    def timed(name, *a):
        if name.endswith('_bytes') or name in ('bignn_abi_version', 'bignn_launch_count'):
            return orig(name, *a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(name, *a)
        e1.record()
        key = name                  # (the key only: holding the argument tensors would pin every activation)
        if name == 'bignn_spmm_f32':
            key = 'bignn_spmm_f32[mode={},D={},rows={}]'.format(a[8], a[7], a[6])
        elif name == 'bignn_gemm_f32':
            key = 'bignn_gemm_f32[ta={},tb={},M={},N={},K={}]'.format(*a[:5])
        elif name in ('bignn_gemm_tc_f32', 'bignn_gemm_tc_masked_f32'):
            key = '{}[M={},N={},K={}]'.format(name, *a[:3])
        elif name == 'bignn_dw_tc_f32':
            key = 'bignn_dw_tc_f32[M={},Np={},Nq={}]'.format(*a[:3])
        elif name == 'bignn_gin_layer_fwd':
            key = 'bignn_gin_layer_fwd[rows={},Din={},Dout={}]'.format(*a[:3])
        rec.append((key, e0, e1, algorithmic_bytes(name, a)))
        return r
    sb = eng.last_static_batch
    n0 = lib.launch_count()
    eng._device_step(sb)           # eager warm-up of this path
    launches_per_step = lib.launch_count() - n0
    if on_start is not None:
        on_start()
    lib.call = timed
    try:
        for _ in range(steps):
            eng._device_step(sb)
        torch.cuda.synchronize()
    finally:
        lib.call = orig
    agg = {}
    for key, e0, e1, nb in rec:
        d = agg.setdefault(key, [0.0, 0, 0.0, True])
        d[0] += e0.elapsed_time(e1)
        d[1] += 1
        if nb is None:
            d[3] = False
        else:
            d[2] += nb
    total = sum(v[0] for v in agg.values())
    rows = sorted(agg.items(), key=lambda kv: -kv[1][0])
    top = []
    for k, (ms, calls, nbytes, known) in rows[:14]:
        e = dict(entry=k, ms_per_step=round(ms / steps, 5), calls_per_step=calls // steps,
                 ms_per_call=round(ms / calls, 5), share=round(ms / total, 4))
        if known and nbytes > 0:
            e['algorithmic_bytes_per_call'] = nbytes / calls
            e['achieved_gbs'] = round(nbytes / (ms * 1e-3) / 1e9, 1)
        top.append(e)
    return dict(total_ms_per_step=total / steps, launches_per_step=launches_per_step, top=top,
                note='eager pass with CUDA events around every C-ABI call (not the CUDA-graph replay that `value` times)')


# DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) over the algorithmic bytes of the same launch, from
# the committed `ncu --set full` captures (profiles/r2_summary.md); a run without a profiler cannot read DRAM counters, so
# `roofline.traffic` = this measured ratio x the algorithmic bytes of the launch being reported
NCU_TRAFFIC = [
    ('bignn_gin_layer_fwd', 0.999,
     'profiles/r2_summary.md: ncu --set full of k_gin_layer_fwd<896,..> at 6 000 245 rows x 64 keeping z and t: dram read '
     '1.612 GB + write 4.602 GB = 6.214 GB = 0.999 x algorithmic; scaled to this launch'),
    ('bignn_dw_tc_f32', 1.003,
     'profiles/r2_summary.md: ncu --set full of k_dw_tc_ring at 6 000 000 rows: dram read 3.073 GB + write 0.009 GB = '
     '1.003 x algorithmic; scaled to this launch'),
]


def step_roofline(prof, peaks):
    """`roofline` of the dominant kernel of the TIMED step: the C-ABI entry point with the largest share of the step's
    device time SUMMED OVER ITS CALLS (the first lower layer has 49 input columns, the others 64: one kernel, two shapes
    -- grouped the way the ncu launch list groups them); achieved = its algorithmic bytes / its time over those calls
    (CUDA events in the eager pass)."""
    peak = peaks.get('hbm_gbs', 6650.0)
    groups = {}
    for e in prof.get('top', []):
        if 'achieved_gbs' not in e:
            continue
        g = groups.setdefault(e['entry'].split('[')[0], dict(ms=0.0, nbytes=0.0, calls=0, share=0.0, shapes=[]))
        g['ms'] += e['ms_per_call'] * e['calls_per_step']
        g['nbytes'] += e['algorithmic_bytes_per_call'] * e['calls_per_step']
        g['calls'] += e['calls_per_step']
        g['share'] += e['share']
        g['shapes'].append(e['entry'])
    if not groups:
        return None
    name, g = max(groups.items(), key=lambda kv: kv[1]['share'])
    ach = g['nbytes'] / (g['ms'] * 1e-3) / 1e9
    traffic, src = None, ('dram bytes of this kernel: see the ncu --set full summary under profiles/ '
                          '(not measurable inside an un-profiled run)')
    for prefix, ratio, cite in NCU_TRAFFIC:
        if name.startswith(prefix):
            traffic = ratio * g['nbytes'] / g['calls']
            src = cite
    out = dict(kernel=name, shapes=g['shapes'], bound='hbm', achieved=round(ach, 1), peak=peak, unit='GB/s',
               frac=round(ach / peak, 4), traffic=traffic, share_of_step=round(g['share'], 4),
               ms_per_launch=round(g['ms'] / g['calls'], 5), launches_per_step=g['calls'],
               algorithmic_bytes=g['nbytes'] / g['calls'],
               peak_source='MEASURED_PEAKS.json hbm_gbs (measured copy peak)' if 'hbm_gbs' in peaks
               else 'fallback 6650 GB/s', traffic_note=src)
    return out


def l2_gather_line(prof, eng):
    """The upper-level SpMM is not an HBM kernel: its feature matrix (drugs x 64 x 4 B) is L2-resident and every
    neighbour is a 256-byte gather out of L2.  Reported beside `roofline`: gathered bytes per launch / launch time."""
    for e in prof.get('top', []):
        if e['entry'].startswith('bignn_spmm_planned_rows_f32') or e['entry'].startswith('bignn_spmm_planned_f32'):
            try:
                nnz = int(eng.data.interaction_combo_nxgraph.csr.nnz)
            except Exception:
                return None
            gathered = 4.0 * 64 * nnz
            return dict(kernel=e['entry'], gathered_bytes_per_launch=gathered, ms_per_launch=e['ms_per_call'],
                        l2_gather_gbs=round(gathered / (e['ms_per_call'] * 1e-3) / 1e9, 1), share_of_step=e['share'],
                        note='L2-resident operand: bound by L2 gather bandwidth, not by HBM (its algorithmic HBM bytes '
                             'are the index stream and one pass over the drugs x 64 matrix)')
    return None


def main():
    args = parse()
    rank = int(os.environ.get('RANK', 0))
    if args.impl == 'reference':
        if rank != 0:
            return
        r = run_reference(args, device=args.ref_device)
        cfg = workload_config(args.workload)
        out = dict(metric=METRIC, value=r['value'], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                   ms_per_step=r['ms_per_step'], higher_is_better=True, scaling='strong', vs_baseline=None, dtype='f32',
                   data='synthetic', impl='reference', config=cfg,
                   cpu_baseline=dict(value=r['value'], unit=UNIT, cores=r['cores'], kind='port', sample=r['sample'],
                                     measured_sample_value=r['measured_sample_value'],
                                     extrapolation_factor=r['extrapolation_factor'], device=args.ref_device),
                   e2e=dict(value=r['value'], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(out))
        return
    res = run_lower_only(args) if args.workload.startswith('mol_1m') else run_ours(args)
    if res is None:
        return
    out, (torch, B, peaks, dev, eng, data, model) = res
    if args.gpus > 1:
        # no destroy_process_group() here: tearing the NCCL communicator down while captured CUDA graphs
        # still hold its kernels blocks forever (measured: a 900 s hang); the process simply exits
        out['gpu_launches'] = None if not hasattr(eng, 'launches_per_step') else eng.launches_per_step * args.steps
        print(json.dumps(out))
        return
    out['gpu_launches'] = int(getattr(eng, 'launches_per_step', 0) * args.steps)
    prof = None
    try:
        if getattr(eng, '_graphs', None):
            # the timed region is over: release the captured graphs' private memory pools before the eager profiling
            # pass (at 20 M atoms the graph pool and an eager step do not fit side by side in 180 GB)
            import gc
            eng._graphs.clear()
            gc.collect()
            torch.cuda.empty_cache()
        prof = kernel_profile(torch, B, eng)
        if not eng.use_cuda_graph:
            out['gpu_launches'] = int(prof['launches_per_step'] * args.steps)
        out['kernel_profile'] = prof
    except Exception as e:          # the headline numbers above are already measured: report, do not lose them
        out['kernel_profile'] = dict(error=repr(e)[:300])
    if not args.skip_rooflines:
        if prof is not None:
            out['roofline'] = step_roofline(prof, peaks)
            try:
                g = l2_gather_line(prof, eng)
                if g is not None:
                    out['upper_spmm_l2_gather'] = g
            except Exception as e:
                out['upper_spmm_l2_gather'] = dict(error=repr(e)[:200])
        # the kernel BASELINE's metric names (segment SpMM) on a > L2 molecule-like graph, stand-alone
        out['roofline_segment_spmm'] = spmm_roofline(torch, B, args.spmm_rows, peaks, dev)
    if not args.skip_gpu_eager:
        try:
            r = run_reference(argparse.Namespace(**vars(args)), sample_steps=3, device=dev)
            out['gpu_eager_baseline'] = dict(value=r['value'], unit=UNIT, kind='port on cuda (torch eager + index_add)',
                                             sample=r['sample'], ms_per_step=r['ms_per_step'])
        except Exception as e:
            out['gpu_eager_baseline'] = dict(error=repr(e)[:300])
    if not args.skip_cpu:
        r = run_reference(argparse.Namespace(**vars(args)), sample_steps=args.cpu_sample_steps)
        out['cpu_baseline'] = dict(value=r['value'], unit=UNIT, cores=r['cores'], kind='port', sample=r['sample'],
                                   ms_per_step=r['ms_per_step'], measured_sample_value=r['measured_sample_value'],
                                   extrapolation_factor=r['extrapolation_factor'])
    print(json.dumps(out))


if __name__ == '__main__':
    main()
