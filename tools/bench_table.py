#!/usr/bin/env python
"""Print markdown tables from bench.py JSON lines (for DESIGN.md):  tools/bench_table.py gpurun_out/x_n1.json [x_n2.json ...]"""
import json
import sys


def line(path):
    for l in reversed(open(path).read().strip().splitlines()):
        if l.startswith("{"):
            return json.loads(l)
    raise SystemExit(f"no JSON line in {path}")


for p in sys.argv[1:]:
    d = line(p)
    T = d["config"].get("timesteps_per_step", 100)
    rl = d["roofline"]
    print(f"\n### {p}: N = {d['n_gpus']}")
    print("| item | value |\n|---|---|")
    print(f"| value | {d['value']:.4g} events/s ({d['ms_per_step'] * 1e3 / T:.1f} us / timestep, {d['steps']} x {T} timesteps, graph={d.get('cuda_graph')}) |")
    print(f"| dominant kernel | `{rl['kernel']}` {rl['kernel_ms_avg'] * 1e3:.1f} us, {rl['algorithmic_bytes_per_launch'] / 1e6:.1f} MB algorithmic -> {rl['achieved']:.0f} GB/s = {rl['frac']:.3f} of {rl['peak']:.0f}; traffic {('%.1f MB' % (rl['traffic'] / 1e6)) if rl.get('traffic') else 'null'} |")
    if d.get("reference_cuda_us"):
        print(f"| reference CUDA kernel, same buffers | {d['reference_cuda_us']} us / timestep ({d.get('vs_reference_cuda')}x) |")
    e = d.get("e2e")
    if e:
        print(f"| e2e | {e['value']:.4g} events/s, {e.get('us_per_timestep')} us / timestep; copies only {e.get('copy_only_us_per_timestep')} us ({e.get('host_link_GBps_all_ranks')} GB/s all ranks); sync {e.get('sync_value')} |")
    if d.get("cpu_baseline"):
        print(f"| cpu_baseline | {d['cpu_baseline']['value']:.3g} events/s, {d['cpu_baseline']['cores']} core, {d['cpu_baseline']['kind']} |")
    if d.get("parity"):
        q = d["parity"]
        print(f"| parity | ok={q['ok']} max_abs_err={q['max_abs_err']:.2g} ids_equal_nccl={q['ids_equal_nccl']} oracle_slice={q['oracle_slice']['max_abs_err']:.2g} |")
    s = d.get("strong_scaling")
    if s:
        print(f"| strong scaling (10 M x 1000) | {s.get('value', 0):.4g} events/s, {s.get('us_per_timestep')} us / timestep{' (= main line)' if s.get('same_as_main_line') else ''} |")
    print(f"| clocks | {d.get('clocks')} |")
    ow = d.get("other_workloads", [])
    if ow:
        print("\n| workload | us / timestep | events/s | kernel: us, frac of HBM peak, traffic MB | reference CUDA us (x) |\n|---|---|---|---|---|")
        for o in ow:
            if "error" in o:
                print(f"| FAILED | {o['error'][:80]} | | | |")
                continue
            r = o["roofline"]
            tr = ("%.1f" % (r["traffic"] / 1e6)) if r.get("traffic") else "null"
            print(f"| {o['workload'][:95]} | {o['ms_per_timestep'] * 1e3:.1f} | {o['value']:.3g} | `{r['kernel']}` {r['kernel_ms_avg'] * 1e3:.1f}, {r['frac']:.3f}, {tr} | {o.get('reference_cuda_us')} ({o.get('vs_reference_cuda')}) |")
