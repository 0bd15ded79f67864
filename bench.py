#!/usr/bin/env python3
"""bench.py - DDC channel x ADC-samples / s on N B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # the CPU golden model on the host cores

Workload (config.workload): BASELINE.json configs[2] at N=1 - 1024 independent DDC channels with random
tuning words over one shared synthetic 12-bit ADC stream, blocks of 2^20 samples - and configs[3] at N=8
(8192 channels sharded by channel, 1024 per GPU, the ADC block broadcast from rank 0 with NCCL): weak scaling.
A step is one ADC block through the whole FPGA receive chain (NCO, mixer, CIC, compensator FIR, Hilbert FIR,
Q delay, 8-byte frames) for every channel of the rank.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CH_PER_GPU = 1024
BLOCK = 1 << 20
A_INT_OPS = 41          # literal HDL-faithful INT32 ops per channel x ADC-sample
SEED = 20261018


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--channels-per-gpu", type=int, default=CH_PER_GPU)
    ap.add_argument("--block", type=int, default=BLOCK)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="ddc", choices=["ddc", "full_chain"],
                    help="ddc (default, the contract's metric: BASELINE configs[2]/[3]) or full_chain (configs[4]: DDC + "
                         "processRxAudio + FFT_doFFT per channel, modes LSB/USB/CW_U/AM/NFM round-robin, DNR + notch on half)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_ddc_rate(n_ch, n_samples, blocks, threads, seed=SEED, warm=0):
    """Golden model (oracle/, C, scalar per channel) on `threads` host threads -> channel*samples/s."""
    from oracle import pyoracle
    import ua3reo_loader
    synth = ua3reo_loader.load().synth
    adc = synth.synth_adc(n_samples, seed=seed)
    bank = pyoracle.GoldenBank(synth.random_fcw(n_ch, seed))
    for _ in range(warm):
        bank.push(adc, threads)
    t = 0.0
    for _ in range(blocks):
        t += bank.push(adc, threads)
    return n_ch * n_samples * blocks / t, t


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation (the golden restatement of the HDL;
    the FPGA design cannot be compiled or simulated here) on all host cores.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    n_ch = 2 * cores
    n_samples = 1 << 18            # bounded sample: 2*cores channels x 2^18 ADC samples per step
    rate, t = cpu_ddc_rate(n_ch, n_samples, args.steps, cores, warm=args.warmup)
    sample = "%d channels x 2^18 ADC samples per step, %d threads, golden C model (oracle/ddc_golden.c)" % (n_ch, cores)
    line = {
        "impl": "reference", "metric": "ddc_channel_adc_samples_per_s", "value": rate, "unit": "channel*samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": "full FPGA RX DDC, random tuning words, one shared 12-bit ADC stream (CPU sample: %s)" % sample,
                   "channels_per_gpu": args.channels_per_gpu, "block_samples": args.block},
        "cpu_baseline": {"value": rate, "unit": "channel*samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "channel*samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import ua3reo_loader
    pkg = ua3reo_loader.load()
    synth = pkg.synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - this repo has no CPU fallback (use --impl reference for the CPU model)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_ch, block, K, W = args.channels_per_gpu, args.block, args.steps, args.warmup

    rx = pkg.Receiver(n_ch, block, device=local)
    fcw_all = synth.random_fcw(n_ch * world, SEED)
    rx.set_fcw(fcw_all[rank * n_ch:(rank + 1) * n_ch])          # channels sharded by rank, contiguous slabs
    ext = torch.cuda.ExternalStream(rx.stream(), device=local)
    full = args.workload == "full_chain"
    audio_host = spec_host = None
    if full:
        # the STM32 stage for every channel; it runs on its own stream one push behind the DDC
        rx.rx_enable(True)
        mix = [(0, 2700), (1, 2700), (4, 500), (10, 6000), (8, 15000)]
        rx.rx_set([rx.rx_defaults(mode=mix[c % 5][0], filter_width=mix[c % 5][1], dnr=(c // 5) % 2, notch=(c // 5) % 2)
                   for c in range(n_ch)])
        nb_max, nf_max = block // 1024 // 192 + 2, block // 1024 // 512 + 2
        audio_host = [torch.empty((n_ch, nb_max, 384), dtype=torch.int32).pin_memory() for _ in range(2)]
        spec_host = [torch.empty((n_ch, nf_max, 256), dtype=torch.float32).pin_memory() for _ in range(2)]
    # N>1: the spectra are gathered back to rank 0 over NVLink (north_star): a device-to-device read of the last push's
    # spectra on the library's copy stream, then an NCCL gather enqueued behind it on the same stream
    gather_spectra = None
    if full and world > 1:
        nf_step = block // 1024 // 512
        assert nf_step * 512 * 1024 == block, "--block must be a multiple of 512 frames for the spectra gather"
        GS = 4                                         # slots: the NCCL kernel only finds a free SM now and then (the front kernel and
        #                                                rx_audio hold them all), so the gather may trail the compute by a step or two
        spec_dev = [torch.empty((n_ch, nf_step, 256), dtype=torch.float32, device="cuda") for _ in range(GS)]
        spec_all = [torch.empty((world, n_ch, nf_step, 256), dtype=torch.float32, device="cuda") for _ in range(GS)] if rank == 0 else None
        spec_all_host = [torch.empty((world, n_ch, nf_step, 256), dtype=torch.float32).pin_memory() for _ in range(GS)] if rank == 0 else None
        cstream = torch.cuda.ExternalStream(rx.copy_stream(), device=local)
        gstream = torch.cuda.Stream()                  # the gather runs here, so that the other pipelined reads are not held up
        gather_pg = dist.new_group()                   # its own communicator: NCCL orders the collectives of one communicator, and a
        #                                                gather that trails the compute must not sit in front of the next broadcast
        ev_spec = [torch.cuda.Event() for _ in range(GS)]
        ev_gdone = [torch.cuda.Event() for _ in range(GS)]
        g_used = [False] * GS

        def gather_spectra(i, to_host):
            b = i % GS
            if g_used[b]:
                cstream.wait_event(ev_gdone[b])        # the gather of step i-2 has read spec_dev[b]
            nf = rx.read_spectra_async(spec_dev[b])
            assert nf == nf_step
            ev_spec[b].record(cstream)
            gstream.wait_event(ev_spec[b])
            with torch.cuda.stream(gstream):
                dist.gather(spec_dev[b], list(spec_all[b].unbind(0)) if rank == 0 else None, dst=0, group=gather_pg)
                if to_host and rank == 0:
                    spec_all_host[b].copy_(spec_all[b], non_blocking=True)
                ev_gdone[b].record(gstream)
            g_used[b] = True

    # synthetic ADC: NB distinct blocks generated once on the host (pinned); rank 0 is the ingest rank
    NB = 4
    host_blocks = torch.from_numpy(synth.synth_adc(NB * block, SEED).reshape(NB, block)).pin_memory()
    dev_blocks = host_blocks.cuda(non_blocking=False) if rank == 0 else None
    bcast = [torch.empty(block, dtype=torch.int16, device="cuda") for _ in range(2)]
    frames_host = [torch.empty((n_ch, block // 1024, 8), dtype=torch.uint8).pin_memory() for _ in range(2)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N>1: rank 0 broadcasts every ADC block over NVLink (NCCL).  The broadcast of block i+1 runs on a side stream
    # while the context's stream computes block i (two buffers, events both ways).
    side = torch.cuda.Stream() if world > 1 else None
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    pending = {}

    def prefetch(i, src_blocks):
        b = i & 1
        with torch.cuda.stream(side):
            if i >= 2:
                side.wait_event(ev_free[b])                      # the push that last read this buffer has finished
            if rank == 0:
                bcast[b].copy_(src_blocks[i % NB], non_blocking=True)
            dist.broadcast(bcast[b].view(torch.uint8), src=0)    # NCCL has no int16: move the bytes
            ev_ready[b].record(side)
        pending[i] = True

    def run_steps(n, src_blocks, after_push=None, to_host=False):
        if world > 1:
            pending.clear()
            prefetch(0, src_blocks)
        for i in range(n):
            if world > 1:
                if i + 1 < n:
                    prefetch(i + 1, src_blocks)
                with torch.cuda.stream(ext):
                    ext.wait_event(ev_ready[i & 1])
                    rx.push(bcast[i & 1])
                    ev_free[i & 1].record(ext)
                if gather_spectra is not None:
                    gather_spectra(i, to_host)
            else:
                with torch.cuda.stream(ext):
                    rx.push(src_blocks[i % NB])
            if after_push is not None:
                after_push()

    def step_device_n(n):
        """n ADC blocks, inputs resident in HBM (rank 0's device copy)."""
        run_steps(n, dev_blocks)

    host_enqueue_ms = [0.0]

    def step_e2e_n(n):
        """Same steps through the host-facing C ABI: every step copies the ADC block from pinned host memory
        (H2D inside the timed region; at N>1 rank 0 ingests and the others receive the broadcast) and reads every
        frame back to pinned host memory (D2H).  The copy of step i is enqueued asynchronously
        (ua3reo_ddc_read_frames_async) so that it overlaps the kernels of step i+1; the final sync is inside the
        timed region."""
        k = [0]

        def pull():
            rx.read_frames_async(frames_host[k[0] & 1])
            if full:                                   # pipelined reads of the STM32 results of this push
                rx.read_audio_async(audio_host[k[0] & 1])
                if gather_spectra is None:             # (N>1: the spectra go to rank 0 through the gather instead)
                    rx.read_spectra_async(spec_host[k[0] & 1])
            k[0] += 1
        t_host = time.perf_counter()
        run_steps(n, host_blocks, after_push=pull, to_host=True)
        host_enqueue_ms[0] = 1e3 * (time.perf_counter() - t_host) / max(n, 1)     # host time to enqueue one step
        rx.sync()
        if gather_spectra is not None:
            gstream.synchronize()

    step_device_n(W)
    barrier()
    int32_peak = pkg.measure_int32_peak(local)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # inside the timed region only the dominant kernel is bracketed by CUDA events (two records per step); the per-kernel
    # breakdown comes from a short second pass afterwards, outside the timed region
    rx.profile_begin(K, kernel="front")
    launches0 = rx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(ext):
        e0.record()
    step_device_n(K)
    if full:
        rx.sync()                                      # the STM32 stage of the last push finishes on its own stream
        if gather_spectra is not None:
            ext.wait_stream(gstream)                   # ... and the last gather
    with torch.cuda.stream(ext):
        e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = rx.launch_count() - launches0
    kms_front, nblocks = rx.profile_end()
    n_break = min(K, 16)
    rx.profile_begin(n_break)
    step_device_n(n_break)
    if full:
        rx.sync()
    barrier()
    kms_all, n_all = rx.profile_end()
    clocks = sampler.stop() if rank == 0 else None

    # end-to-end through the C ABI with host buffers
    step_e2e_n(max(3, W))
    barrier()
    with torch.cuda.stream(ext):
        e0.record()
    step_e2e_n(K)
    with torch.cuda.stream(ext):
        e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    units = float(n_ch) * world * block * K
    value = units / (ms * 1e-3)
    e2e = units / (ms_e2e * 1e-3)
    front_s = kms_front["front"] * 1e-3 / max(nblocks, 1)
    achieved = A_INT_OPS * float(n_ch) * block / front_s if front_s > 0 else 0.0

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = host_cores()
            c_ch, c_n = 4 * cores, 1 << 19
            blocks = 1
            rate, t = cpu_ddc_rate(c_ch, c_n, 1, cores)
            while t < 10.0 and blocks < 64:                       # bounded: about 10-20 s of CPU work
                r2, t2 = cpu_ddc_rate(c_ch, c_n, min(blocks * 2, 64), cores)
                rate, t, blocks = r2, t + t2, min(blocks * 2, 64)
            cpu = {"value": rate, "unit": "channel*samples/s", "cores": cores, "kind": "port",
                   "sample": "%d channels x 2^19 ADC samples x %d blocks, %d threads, golden C model "
                             "(oracle/ddc_golden.c; the FPGA HDL itself cannot be built here)" % (c_ch, blocks, cores)}
        variant = os.environ.get("UA3REO_FRONT_VARIANT", "0")
        front_name = "ddc_front_kernel" if (variant == "1" or n_ch < 256) else "ddc_front_bt_kernel"
        static = {}
        try:
            static = json.load(open(os.path.join(ROOT, "profiles", "r01_front_kernel_static.json"))).get(front_name, {})
        except Exception:
            pass
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        step_bytes = 2.0 * block * (n_ch // 32) + 2 * 80.0 * n_ch * (block // 512) + 8.0 * n_ch * (block // 1024)
        d2h = n_ch * (block // 1024) * 8
        if full:
            d2h += int(n_ch * (block / 1024.0 / 192.0) * 384 * 4 + n_ch * (block / 1024.0 / 512.0) * 256 * 4)
        workload = ("BASELINE configs[2]/[3]: %d independent DDC channels per GPU (random tuning words, seed %d) "
                    "over one shared synthetic 12-bit ADC stream, blocks of %d samples; full FPGA RX chain "
                    "(NCO+mixer+CIC/512+compensator FIR+Hilbert FIR+Q delay+8-byte frames); "
                    "N>1: channels sharded by rank, ADC block NCCL-broadcast from rank 0" % (n_ch, SEED, block))
        if full:
            workload = ("BASELINE configs[4]: full per-channel RX chain for %d channels per GPU - " % n_ch) + workload.split(": ", 1)[1] + \
                       "; then processRxAudio + FFT_doFFT per channel (modes LSB/USB/CW_U/AM/NFM round-robin, DNR + notch on half)" + \
                       ("; spectra NCCL-gathered to rank 0 every step" if world > 1 else "")
        line = {
            "metric": "rx_chain_channel_adc_samples_per_s" if full else "ddc_channel_adc_samples_per_s", "value": value,
            "unit": "channel*samples/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64+f32" if full else "int64", "data": "synthetic",
            "config": {"workload": workload,
                       "channels_total": n_ch * world, "channels_per_gpu": n_ch, "block_samples": block,
                       "real_time_channels": value / 49152000.0,
                       "l2": "per-step working set ~%.0f MB (chunk records + frames) exceeds the 126 MB L2; no flush needed"
                             % (step_bytes / 1e6)},
            "e2e": {"value": e2e, "unit": "channel*samples/s", "h2d_bytes_per_step": 2 * block,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / K, "host_enqueue_ms_per_step": host_enqueue_ms[0]},
            "gpu_launches": int(launches),
            "roofline": {"bound": "int32_alu", "kernel": front_name, "achieved": achieved / 1e12, "peak": int32_peak / 1e12,
                         "unit": "TOP/s (INT32)", "frac": (achieved / int32_peak) if int32_peak else None,
                         "traffic": static.get("traffic_bytes_per_launch") if (n_ch, block) == (1024, 1 << 20) else None,
                         "executed_ops_per_unit": static.get("sass_instructions_per_unit"),
                         "issue_frac": (static["sass_instructions_per_unit"] * float(n_ch) * block / front_s / int32_peak)
                                       if static.get("sass_instructions_per_unit") and front_s > 0 and int32_peak else None,
                         "note": "achieved counts the 41 algorithmic INT32 ops per channel-sample of the HDL-literal formulation "
                                 "(SURVEY.md 8d), which is why frac can exceed 1: the kernel executes executed_ops_per_unit "
                                 "of them (table-driven NCO output stage, shifts fused into adds); issue_frac = executed ops / peak",
                         "ops_per_unit": A_INT_OPS, "units_per_launch": float(n_ch) * block,
                         "kernel_ms": front_s * 1e3,
                         "kernel_share_of_step": front_s * 1e3 / (ms / K),
                         "all_kernels_ms_per_step": {k: v / max(n_all, 1) for k, v in kms_all.items()},
                         "all_kernels_note": "per-kernel times from a separate pass of %d steps right after the timed region "
                                             "(events around every kernel); kernel_ms is from the timed region itself" % n_break,
                         "peak_source": "ua3reo_measure_int32_peak (IMAD+LOP3+IADD3 chains), measured live on this GPU; "
                                        "MEASURED_PEAKS.json has no integer peak",
                         "hbm": {"algorithmic_gbs": step_bytes / (ms / K * 1e-3) / 1e9,
                                 "peak_gbs": peaks.get("hbm_gbs"), "note": "HBM is not the bound (SURVEY.md 8d)"}},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line))
    # tensors that were used on the library's copy stream must be released while that stream still exists (the caching
    # allocator records an event on every stream a block was used on when the block is freed)
    if gather_spectra is not None:
        import gc
        torch.cuda.synchronize()
        del gather_spectra, spec_dev, spec_all, spec_all_host, cstream, gstream, ev_spec, ev_gdone, gather_pg
        gc.collect()
        torch.cuda.synchronize()
    rx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
