#!/usr/bin/env python3
"""bench.py - DDC channel x ADC-samples / s on N B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU implementation on the host cores

The default run measures two workloads back to back and prints ONE line:
  * headline (top-level keys): BASELINE.json configs[2] at N=1 - 1024 independent DDC channels with random tuning words
    over one shared synthetic 12-bit ADC stream, blocks of 2^20 samples - and configs[3] at N=8 (8192 channels sharded by
    channel, 1024 per GPU, the ADC block broadcast from rank 0 with NCCL): weak scaling.  A step is one ADC block through
    the whole FPGA receive chain (NCO, mixer, CIC, compensator FIR, Hilbert FIR, Q delay, 8-byte frames) for every
    channel of the rank.  `--scaling strong` fixes the total at --total-channels (8192) instead.
  * "full_chain" (nested object, same keys): BASELINE.json configs[4], the north_star target - 4096 channels per GPU
    through the DDC AND the STM32 stage (processRxAudio + FFT_doFFT per channel: modes LSB/USB/CW_U/AM/NFM round-robin, DNR
    + notch on half), spectra gathered to rank 0 at N>1, plus the TX DUC ("tx_duc").  Its cpu_baseline is the reference
    firmware's own C (oracle/_ref/fw_rx, kind "reference") behind the golden DDC.
Every workload adds: "parity" (after the timed region every rank pushes one more block through the same path - the
NCCL-broadcast one at N>1 - and compares two of its channels bit for bit with the golden model, the run FAILS otherwise),
"sustained" (>= 3 s of back-to-back steps with the clocks sampled) and the per-kernel breakdown.
`--workload ddc|full_chain` runs one of the two alone (full_chain then becomes the top-level line).
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
FULL_CH_PER_GPU = 4096
BLOCK = 1 << 20
A_INT_OPS = 41          # literal HDL-faithful INT32 ops per channel x ADC-sample
SEED = 20261018
MODE_MIX = [(0, 2700), (1, 2700), (4, 500), (10, 6000), (8, 15000)]     # (TRX_MODE_*, Filter_Width): LSB USB CW_U AM NFM


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--channels-per-gpu", type=int, default=None)
    ap.add_argument("--full-channels-per-gpu", type=int, default=FULL_CH_PER_GPU)
    ap.add_argument("--block", type=int, default=BLOCK)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--sustained-seconds", type=float, default=3.0)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--total-channels", type=int, default=8192, help="--scaling strong: channels over all GPUs (configs[3])")
    ap.add_argument("--workload", default="all", choices=["all", "ddc", "full_chain"])
    ap.add_argument("--comm-sms", type=int, default=None,
                    help="N>1: SMs the front kernel leaves to the NCCL broadcast kernel (default 1; NCCL is held to as many channels)")
    ap.add_argument("--adc-transport", default="auto", choices=["auto", "ipc", "nccl"],
                    help="N>1: how the ADC block reaches the ranks: ipc = ua3reo_fanout_* (copy engines over CUDA IPC mappings, "
                         "no SM), nccl = NCCL broadcast on a side stream (one SM set aside); auto = ipc, nccl if it cannot be set up")
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
        return self

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
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------------------------------------
# CPU legs (the only places that execute oracle/): the golden DDC model, the host-built firmware, the translated HDL
# ----------------------------------------------------------------------------------------------------------------------
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


def cpu_ddc_bounded(cores, budget_s=10.0):
    c_ch, c_n = 4 * cores, 1 << 19
    blocks = 1
    rate, t = cpu_ddc_rate(c_ch, c_n, 1, cores)
    while t < budget_s and blocks < 64:                       # bounded: about 10-20 s of CPU work
        r2, t2 = cpu_ddc_rate(c_ch, c_n, min(blocks * 2, 64), cores)
        rate, t, blocks = r2, t + t2, min(blocks * 2, 64)
    return rate, "%d channels x 2^19 ADC samples x %d blocks, %d threads, golden C model (oracle/ddc_golden.c, " \
                 "pinned edge for edge to the reference's VHDL by tests/test_hdl_pin.py)" % (c_ch, blocks, cores)


def synth_iq_frames(n, seed=SEED):
    """48 kSPS I/Q: USB two-tone (1.0 kHz + 1.9 kHz above the carrier) + noise as 8-byte frames (SPEC = VOICE); SURVEY 8(d)"""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 48000.0
    z = 6000 * np.exp(2j * np.pi * 1000 * t) + 4000 * np.exp(2j * np.pi * 1900 * t) + rng.normal(0, 30, n) + 1j * rng.normal(0, 30, n)
    i, q = np.rint(z.real).astype(np.int16), np.rint(z.imag).astype(np.int16)
    f = np.zeros((n, 8), np.uint8)
    for w, v in ((0, q), (1, i), (2, q), (3, i)):
        f[:, 2 * w] = (v.view(np.uint16) >> 8).astype(np.uint8)
        f[:, 2 * w + 1] = (v.view(np.uint16) & 0xFF).astype(np.uint8)
    return f


def full_chain_settings(c):
    mode, width = MODE_MIX[c % 5]
    return dict(mode=mode, filter_width=width, dnr=(c // 5) % 2, notch=(c // 5) % 2)


def cpu_fw_rx_rate(cores, budget_s=10.0):
    """The reference firmware's own processRxAudio + FFT_doFFT (oracle/_ref/fw_rx: audio_processor.c, fft.c ... compiled
    unmodified), one process per core, the bench's mode mix -> frames/s over all cores (None when the binary is missing)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle
    if not pyoracle.have_fw_rx():
        return None, "oracle/_ref/fw_rx not built"
    probe = synth_iq_frames(48000 * 2)
    t0 = time.perf_counter()
    pyoracle.run_fw_rx(probe, full_chain_settings(4))
    per_frame = (time.perf_counter() - t0) / probe.shape[0]
    n = int(min(max(budget_s / max(per_frame, 1e-9), 48000 * 4), 48000 * 120))
    frames = synth_iq_frames(n)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(lambda c: pyoracle.run_fw_rx(frames, full_chain_settings(c))["audio"].shape, range(cores)))
    dt = time.perf_counter() - t0
    return cores * n / dt, "%d processes x %.0f s of 48 kSPS I/Q, reference firmware C host-built (oracle/_ref/fw_rx)" % (cores, n / 48000.0)


def cpu_hdl_translation_rate():
    """One channel through the C translation of the reference's VHDL (oracle/_ref/libua3_hdl.so), one core."""
    try:
        from oracle import hdl_ref, pyoracle
        if not os.path.exists(hdl_ref.LIB):
            return None
        n = 1 << 18
        import ua3reo_loader
        adc = ua3reo_loader.load().synth.synth_adc(n, seed=SEED)
        x_i, x_q = pyoracle.golden_mixer(adc, 605867)
        t0 = time.perf_counter()
        hdl_ref.rx_chain(x_i, x_q, t_rx=517)
        return n / (time.perf_counter() - t0)
    except Exception:
        return None


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on all host cores, rank 0 only.
    DDC workload: the golden C model (the FPGA design is VHDL + an encrypted NCO; its filters DO run here, translated
    from their own source by tools/vhdl_eval.py, and tests/test_hdl_pin.py shows the golden model equals them edge for
    edge - the translation is 10x slower, so the faster, identical golden model is the baseline that is timed).
    Full-chain workload: golden DDC + the reference firmware's own C (oracle/_ref/fw_rx)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    n_ch = 2 * cores
    n_samples = 1 << 18            # bounded sample: 2*cores channels x 2^18 ADC samples per step
    rate, t = cpu_ddc_rate(n_ch, n_samples, args.steps, cores, warm=args.warmup)
    sample = "%d channels x 2^18 ADC samples per step, %d threads, golden C model (oracle/ddc_golden.c)" % (n_ch, cores)
    ch_gpu = args.channels_per_gpu or CH_PER_GPU

    def line_for(metric, value, ms_per_step, workload, kind, sample):
        return {
            "impl": "reference", "metric": metric, "value": value, "unit": "channel*samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "int64" if metric.startswith("ddc") else "int64+f32", "data": "synthetic",
            "config": {"workload": workload, "channels_per_gpu": ch_gpu, "block_samples": args.block},
            "cpu_baseline": {"value": value, "unit": "channel*samples/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "channel*samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
    line = line_for("ddc_channel_adc_samples_per_s", rate, 1e3 * t / args.steps,
                    "full FPGA RX DDC, random tuning words, one shared 12-bit ADC stream (CPU sample: %s)" % sample, "port", sample)
    hdl = cpu_hdl_translation_rate()
    if hdl:
        line["cpu_baseline"]["hdl_translation_single_core"] = {
            "value": hdl, "unit": "channel*samples/s",
            "note": "the reference's own VHDL filters translated to C (oracle/_ref/libua3_hdl.so), one channel on one core; "
                    "bit-identical to the golden model, which is what is timed above"}
    if args.workload in ("all", "full_chain"):
        fw, fw_sample = cpu_fw_rx_rate(cores, budget_s=8.0)
        if fw:
            both = 1.0 / (1.0 / rate + 1.0 / (fw * 1024.0))          # the same cores run the DDC, then the firmware stage
            fc = line_for("rx_chain_channel_adc_samples_per_s", both, None,
                          "full per-channel RX chain on the host cores: golden DDC, then the reference firmware's processRxAudio + "
                          "FFT_doFFT (mode mix of the GPU arm)", "reference", sample + "; then " + fw_sample)
            fc["cpu_baseline"]["firmware_stage_frames_per_s"] = fw
            fc["cpu_baseline"]["ddc_stage_channel_samples_per_s"] = rate
            if args.workload == "full_chain":
                line = fc
            else:
                line["full_chain"] = fc
    emit(line)


# ----------------------------------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------------------------------
def measure(args, pkg, torch, dist, workload, n_ch, world, rank, local, want_cpu):
    """One workload on this rank's GPU (all ranks call it); returns the result dict on rank 0, None elsewhere."""
    synth = pkg.synth
    block, K, W = args.block, args.steps, args.warmup
    full = workload == "full_chain"

    rx = pkg.Receiver(n_ch, block, device=local)
    fcw_all = synth.random_fcw(n_ch * world, SEED)
    my_fcw = fcw_all[rank * n_ch:(rank + 1) * n_ch]                      # channels sharded by rank, contiguous slabs
    rx.set_fcw(my_fcw)
    ext = torch.cuda.ExternalStream(rx.stream(), device=local)
    # N>1: rank 0's ADC blocks reach every rank over NVLink, block i+1 while block i is computed.  Preferred: ua3reo_fanout_*
    # (sharding.AdcFanout: the ingest rank's copy engines write into every rank's slot through CUDA IPC mappings, the consumer
    # streams wait on a flag word - no kernel, no SM).  Fallback: an NCCL broadcast on a side stream (sharding.AdcBroadcaster:
    # two buffers, events both ways); its kernel cannot share an SM with a front CTA, so one SM is set aside for it.
    bc, transport = None, "none"
    if world > 1 and args.adc_transport in ("auto", "ipc"):
        try:
            bc = pkg.sharding.AdcFanout(rx.lib, block, local, rx.stream(), src=0, dist=dist)
            transport = "cuda-ipc copy engines (ua3reo_fanout_*)"
        except pkg.UA3Error as e:
            if args.adc_transport == "ipc":
                raise
            if rank == 0:
                print("bench.py: %s - falling back to the NCCL broadcast" % e, file=sys.stderr)
    if world > 1 and bc is None:
        bc = pkg.sharding.AdcBroadcaster(block, torch.device("cuda", local), src=0, dist=dist, consumer_stream=ext)
        transport = "nccl broadcast"
    gather_transport = "nccl gather" if (full and world > 1) else "none"
    audio_host = spec_host = None
    if full:
        # the STM32 stage for every channel; it runs on its own stream one push behind the DDC
        rx.rx_enable(True)
        rx.rx_set([rx.rx_defaults(**full_chain_settings(rank * n_ch + c)) for c in range(n_ch)])
        nb_max, nf_max = block // 1024 // 192 + 2, block // 1024 // 512 + 2
        audio_host = [torch.empty((n_ch, nb_max, 384), dtype=torch.int32).pin_memory() for _ in range(2)]
        spec_host = [torch.empty((n_ch, nf_max, 256), dtype=torch.float32).pin_memory() for _ in range(2)]
    # N>1: the spectra are gathered back to rank 0 over NVLink (north_star): a device-to-device read of the last push's
    # spectra on the library's copy stream, then an NCCL gather enqueued behind it
    gather_spectra = None
    if full and world > 1:
        nf_step = block // 1024 // 512
        assert nf_step * 512 * 1024 == block, "--block must be a multiple of 512 frames for the spectra gather"
        GS = 4                                         # slots: the gather may trail the compute by a step or two
        spec_dev = [torch.empty((n_ch, nf_step, 256), dtype=torch.float32, device="cuda") for _ in range(GS)]
        spec_all = [torch.empty((world, n_ch, nf_step, 256), dtype=torch.float32, device="cuda") for _ in range(GS)] if rank == 0 else None
        spec_all_host = [torch.empty((world, n_ch, nf_step, 256), dtype=torch.float32).pin_memory() for _ in range(GS)] if rank == 0 else None
        cstream = torch.cuda.ExternalStream(rx.copy_stream(), device=local)
        gstream = torch.cuda.Stream()                  # the gather runs here, so that the other pipelined reads are not held up
        gather_pg = dist.new_group()                   # its own communicator: a gather that trails the compute must not queue in
        #                                                front of the next ADC broadcast (NCCL orders one communicator's collectives)
        ev_spec = [torch.cuda.Event() for _ in range(GS)]
        ev_gdone = [torch.cuda.Event() for _ in range(GS)]
        g_used = [False] * GS
        # preferred: ua3reo_gather_* - every rank's copy engine writes its slab into rank 0's buffer (CUDA IPC), no NCCL kernel
        slab_gather, gathered_view = None, [None]
        if transport.startswith("cuda-ipc"):
            try:
                slab_gather = pkg.sharding.SlabGather(rx.lib, n_ch * nf_step * 256 * 4, local, root=0, dist=dist, n_buffers=GS)
                gather_transport = "cuda-ipc copy engines (ua3reo_gather_*)"
            except pkg.UA3Error as e:
                if rank == 0:
                    print("bench.py: %s - the spectra gather stays on NCCL" % e, file=sys.stderr)

        class _Raw:                                    # a device range as a torch tensor (zero copy)
            def __init__(self, ptr, nbytes):
                self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

        def gather_spectra_ce(i, to_host):
            b = i % GS
            nf = rx.read_spectra_async(spec_dev[b])            # device-to-device, on the library's copy stream ...
            assert nf == nf_step
            slab_gather.send(spec_dev[b].data_ptr(), rx.copy_stream())     # ... and so is the send: ordered behind it
            if rank == 0:
                ptr, stride = slab_gather.acquire(gstream.cuda_stream)      # gstream waits for every rank's arrival word
                if to_host:
                    with torch.cuda.stream(gstream):
                        allr = torch.as_tensor(_Raw(ptr, world * stride), device="cuda").view(world, stride)
                        gathered_view[0] = allr[:, :n_ch * nf_step * 256 * 4].contiguous().view(torch.float32).view(world, n_ch, nf_step, 256).cpu()
                slab_gather.release(gstream.cuda_stream)

        def gather_spectra_nccl(i, to_host):
            b = i % GS
            if g_used[b]:
                cstream.wait_event(ev_gdone[b])        # the gather that last used slot b has read spec_dev[b]
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
            if to_host and rank == 0:
                gathered_view[0] = spec_all_host[b].view(world, n_ch, nf_step, 256)

        gather_spectra = gather_spectra_ce if slab_gather is not None else gather_spectra_nccl

    # SMs kept out of the front kernel for NCCL kernels (a front CTA owns its SM): the broadcast, the spectra gather
    comm_sms = args.comm_sms if args.comm_sms is not None else \
        (1 if world > 1 and (transport == "nccl broadcast" or gather_transport == "nccl gather") else 0)
    if world > 1:
        rx.reserve_sms(comm_sms)

    # synthetic ADC: NB distinct blocks generated once on the host (pinned); rank 0 is the ingest rank
    NB = 4
    host_np = synth.synth_adc(NB * block, SEED).reshape(NB, block)
    host_blocks = torch.from_numpy(host_np).pin_memory()
    dev_blocks = host_blocks.cuda(non_blocking=False) if rank == 0 else None
    frames_host = [torch.empty((n_ch, block // 1024, 8), dtype=torch.uint8).pin_memory() for _ in range(2)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(n, src_blocks, after_push=None, to_host=False):
        if bc is not None:
            bc.prefetch(src_blocks[0] if rank == 0 else None)
        for i in range(n):
            if bc is not None:
                if i + 1 < n:
                    bc.prefetch(src_blocks[(i + 1) % NB] if rank == 0 else None)
                buf = bc.acquire()
                rx.push(buf, assume_ordered=True)
                bc.release(buf)
                if gather_spectra is not None:
                    gather_spectra(i, to_host)
            else:
                rx.push(src_blocks[i % NB], assume_ordered=True)       # resident blocks, produced long ago
            if after_push is not None:
                after_push()

    def step_device_n(n):
        """n ADC blocks, inputs resident in HBM (rank 0's device copy)."""
        run_steps(n, dev_blocks)

    host_enqueue_ms = [0.0]

    def step_e2e_n(n):
        """Same steps through the host-facing C ABI: every step copies the ADC block from pinned host memory (H2D inside the
        timed region; at N>1 rank 0 ingests and the others receive the broadcast) and reads every frame back to pinned host
        memory (D2H); full chain: the codec audio words and the spectra instead (the frames never leave the device there).  The copies of step i are enqueued asynchronously so that they
        overlap the kernels of step i+1; the final sync is inside the timed region."""
        k = [0]

        def pull():
            if not full:
                rx.read_frames_async(frames_host[k[0] & 1])
            else:                                      # the frames are consumed on the device by the STM32 stage; what a full-chain
                                                       # user reads are its results: pipelined reads of this push's audio and spectra
                rx.read_audio_async(audio_host[k[0] & 1])
                # every rank sinks its own spectra into host memory over its own PCIe link; at N>1 the NCCL gather
                # additionally assembles all of them in rank 0's HBM (for a consumer on that device)
                rx.read_spectra_async(spec_host[k[0] & 1])
            k[0] += 1
        t_host = time.perf_counter()
        if bc is None:
            for i in range(n):
                rx.push(host_np[i % NB])               # numpy view of the pinned blocks: ua3reo_ddc_push (H2D inside)
                pull()
        else:
            run_steps(n, host_blocks, after_push=pull, to_host=False)
        host_enqueue_ms[0] = 1e3 * (time.perf_counter() - t_host) / max(n, 1)     # host time to enqueue one step
        rx.sync()
        if gather_spectra is not None:
            gstream.synchronize()

    def finish_device():
        if full:
            rx.sync()                                  # the STM32 stage of the last push finishes on its own stream
            if gather_spectra is not None:
                ext.wait_stream(gstream)               # ... and the last gather

    step_device_n(W)
    finish_device()
    barrier()
    int32_peak = pkg.measure_int32_peak(local)
    lds_peak = pkg.measure_lds_peak(local)
    barrier()

    sampler = ClockSampler(local).start() if rank == 0 else None
    # inside the timed region only the dominant kernel is bracketed by CUDA events (two records per step); the per-kernel
    # breakdown comes from a short second pass afterwards, outside the timed region
    rx.profile_begin(K, kernel="front")
    launches0 = rx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(ext)
    step_device_n(K)
    finish_device()
    e1.record(ext)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = rx.launch_count() - launches0
    kms_front, nblocks = rx.profile_end()
    n_break = min(K, 16)
    rx.profile_begin(n_break)
    step_device_n(n_break)
    finish_device()
    barrier()
    kms_all, n_all = rx.profile_end()
    clocks = sampler.stop() if rank == 0 else None

    # end-to-end through the C ABI with host buffers
    step_e2e_n(max(3, W))
    barrier()
    e0.record(ext)
    step_e2e_n(K)
    e1.record(ext)
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    # what the box's host side can take: every rank copies 4 x 64 MB device -> pinned host at the same time (the ceiling of e2e
    # at N>1 is this aggregate, not the GPUs: DESIGN.md 6)
    link_dev = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    link_host = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    link_host.copy_(link_dev, non_blocking=True)
    barrier()
    lk0, lk1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lk0.record()
    for _ in range(4):
        link_host.copy_(link_dev, non_blocking=True)
    lk1.record()
    barrier()
    ms_link = lk0.elapsed_time(lk1)
    del link_dev, link_host

    # sustained: >= args.sustained_seconds of back-to-back steps, clocks sampled
    sustained = None
    if not args.no_sustained:
        ms_all = ms
        if world > 1:                                  # every rank must run the SAME number of steps (one broadcast each)
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_all = float(t[0])
        n_sus = max(K, int(args.sustained_seconds * 1e3 / (ms_all / K)) + 1)
        sus_sampler = ClockSampler(local).start() if rank == 0 else None
        barrier()
        e0.record(ext)
        step_device_n(n_sus)
        finish_device()
        e1.record(ext)
        barrier()
        ms_sus = e0.elapsed_time(e1)
        sus_clocks = sus_sampler.stop() if rank == 0 else None
        sustained = [ms_sus, n_sus, sus_clocks]

    # parity: one more block through the same path, two channels of every rank against the golden model
    from oracle import pyoracle                        # the checker, outside every timed region
    rx.sync()
    rx.reset()
    if bc is not None:
        bc.prefetch(dev_blocks[1] if rank == 0 else None)
        buf = bc.acquire()
        rx.push(buf, assume_ordered=True)
        bc.release(buf)
    else:
        rx.push(dev_blocks[1], assume_ordered=True)
    got = rx.read_frames()
    picks = sorted({0, n_ch - 1})
    ddc_ok = True
    for c in picks:
        ref = pyoracle.GoldenDDC(int(my_fcw[c])).push(host_np[1])
        if not np.array_equal(got[c], ref):
            ddc_ok = False
            bad = np.argwhere(got[c] != ref)
            sys.stderr.write("bench.py: rank %d channel %d (fcw %d): %d of %d frame bytes differ, first at frame %d byte %d; "
                             "frames got %s expected %s\n" % (rank, c, int(my_fcw[c]), bad.shape[0], ref.size, bad[0][0], bad[0][1],
                                                               got.shape, ref.shape))
    audio_ok, audio_note = None, None
    if full:
        if pyoracle.have_fw_rx():
            audio = rx.read_audio()
            spectra = rx.read_spectra()
            audio_ok = True
            worst = 0.0
            for c in picks:
                ref = pyoracle.run_fw_rx(got[c], full_chain_settings(rank * n_ch + c))
                nb, nf = audio.shape[1], spectra.shape[1]
                d = np.abs(audio[c].astype(np.float64) - ref["audio"][:nb])
                pk = max(1.0, float(np.abs(ref["audio"][:nb]).max()))
                ds = np.abs(spectra[c] - ref["spectra"][:nf]).max() if nf else 0.0
                worst = max(worst, float(d.max()) / pk, float(ds))
                audio_ok = audio_ok and float(d.max()) / pk <= 1e-5 and float(ds) <= 1e-5
            audio_note = "max |d|/peak over audio and spectra of the checked channels vs oracle/_ref/fw_rx: %.3g (tolerance 1e-5)" % worst
        else:
            audio_note = "oracle/_ref/fw_rx not present: STM32-stage parity not checked in this run"

    # N>1: the gather itself - rank 0's buffer must hold every rank's spectra of this push (CRC of each slab against the CRC
    # its owner computed from its own host read)
    gather_ok = None
    if gather_spectra is not None:
        import zlib
        mine = np.ascontiguousarray(rx.read_spectra())
        gather_spectra(0, True)
        gstream.synchronize()
        crcs = [None] * world
        dist.all_gather_object(crcs, zlib.crc32(mine.tobytes()))
        if rank == 0:
            got_all = gathered_view[0].numpy()
            gather_ok = all(zlib.crc32(np.ascontiguousarray(got_all[r]).tobytes()) == crcs[r] for r in range(world)) and bool(np.abs(got_all).max() > 0)

    # TX DUC (configs[4] "plus TX DUC interpolation"): n_tx 48 kHz I/Q samples per channel -> n_tx * 1024 DAC words
    tx_duc = None
    if full:
        n_tx = 64
        rx.duc_enable(n_tx)
        iq = np.random.default_rng(SEED).integers(-20000, 20000, (n_ch, n_tx, 2)).astype(np.int16)
        rx.duc_push(iq)
        rx.sync()
        reps = 3
        l0 = rx.launch_count()
        barrier()
        e0.record(ext)
        for _ in range(reps):
            rx.duc_push(iq)
        e1.record(ext)
        barrier()
        ms_duc = e0.elapsed_time(e1)
        tx_duc = [ms_duc, reps, n_tx, rx.launch_count() - l0]

    red = [ms, ms_e2e, sustained[0] if sustained else 0.0, tx_duc[0] if tx_duc else 0.0, ms_link]
    oks = [1.0 if ddc_ok else 0.0, 1.0 if audio_ok else 0.0]
    if world > 1:
        t = torch.tensor(red, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        red = [float(v) for v in t]
        t = torch.tensor(oks, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        oks = [float(v) for v in t]
    ms, ms_e2e = red[0], red[1]

    units = float(n_ch) * world * block * K
    value = units / (ms * 1e-3)
    e2e = units / (ms_e2e * 1e-3)
    front_s = kms_front["front"] * 1e-3 / max(nblocks, 1)
    achieved = A_INT_OPS * float(n_ch) * block / front_s if front_s > 0 else 0.0
    line = None
    if rank == 0:
        cores = host_cores()
        cpu = None
        if want_cpu:
            rate, sample = cpu_ddc_bounded(cores)
            cpu = {"value": rate, "unit": "channel*samples/s", "cores": cores, "kind": "port", "sample": sample}
            if full:
                fw, fw_sample = cpu_fw_rx_rate(cores)
                if fw:
                    cpu = {"value": 1.0 / (1.0 / rate + 1.0 / (fw * 1024.0)), "unit": "channel*samples/s", "cores": cores,
                           "kind": "reference", "sample": sample + "; then " + fw_sample,
                           "ddc_stage_channel_samples_per_s": rate, "firmware_stage_frames_per_s": fw,
                           "note": "the firmware stage is the reference's own C (audio_processor.c, fft.c ... host-built unmodified); "
                                   "the DDC in front of it is the golden model, the same cores run both"}
            else:
                hdl = cpu_hdl_translation_rate()
                if hdl:
                    cpu["hdl_translation_single_core"] = {"value": hdl, "unit": "channel*samples/s",
                                                          "note": "the reference's VHDL filters translated to C (oracle/_ref/libua3_hdl.so), one core"}
        variant = os.environ.get("UA3REO_FRONT_VARIANT", "0")
        front_name = "ddc_front_kernel" if (variant == "1" or (n_ch < 256 and variant not in ("2", "3"))) else \
            ("ddc_front_bt_kernel" if variant == "2" else "ddc_front_tc_kernel")
        static = {}
        try:
            static = json.load(open(os.path.join(ROOT, "profiles", "front_kernel_static.json"))).get(front_name, {})
        except Exception:
            pass
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        rec_bytes = 80.0
        step_bytes = 2.0 * block * (n_ch // 32) + 2 * rec_bytes * n_ch * (block // 512) + 8.0 * n_ch * (block // 1024)
        d2h = n_ch * (block // 1024) * 8
        if full:
            d2h = int(n_ch * (block / 1024.0 / 192.0) * 384 * 4 + n_ch * (block / 1024.0 / 512.0) * 256 * 4)
        workload_s = ("BASELINE configs[2]/[3]: %d independent DDC channels per GPU (random tuning words, seed %d) "
                      "over one shared synthetic 12-bit ADC stream, blocks of %d samples; full FPGA RX chain "
                      "(NCO+mixer+CIC/512+compensator FIR+Hilbert FIR+Q delay+8-byte frames); "
                      "N>1: channels sharded by rank, ADC block sent from rank 0 (config.adc_transport)" % (n_ch, SEED, block))
        if full:
            workload_s = ("BASELINE configs[4]: full per-channel RX chain for %d channels per GPU - " % n_ch) + workload_s.split(": ", 1)[1] + \
                         "; then processRxAudio + FFT_doFFT per channel (modes LSB/USB/CW_U/AM/NFM round-robin, DNR + notch on half)" + \
                         ("; spectra NCCL-gathered into rank 0's HBM every step, host copies per rank" if world > 1 else "")
        issue = (static["sass_instructions_per_unit"] * float(n_ch) * block / front_s / int32_peak) \
            if static.get("sass_instructions_per_unit") and front_s > 0 and int32_peak else None
        units = float(n_ch) * block
        common = {"kernel": front_name, "units_per_launch": units, "kernel_ms": front_s * 1e3,
                  "kernel_share_of_step": front_s * 1e3 / (ms / K),
                  "traffic": static.get("traffic_bytes_per_launch") if (n_ch, block) == (1024, 1 << 20) else None,
                  "traffic_ratio": (static["traffic_bytes_per_launch"] / (2.0 * block + 8.0 * n_ch * (block // 1024))
                                    if static.get("traffic_bytes_per_launch") and (n_ch, block) == (1024, 1 << 20) else None),
                  "executed_instructions_per_unit": static.get("sass_instructions_per_unit"),
                  "issue_frac": issue, "static_source": static.get("source"),
                  "all_kernels_ms_per_step": {k: v / max(n_all, 1) for k, v in kms_all.items()},
                  "all_kernels_note": "per-kernel times from a separate pass of %d steps right after the timed region "
                                      "(events around every kernel); kernel_ms is from the timed region itself" % n_break,
                  "hbm": {"algorithmic_gbs": step_bytes / (ms / K * 1e-3) / 1e9,
                          "peak_gbs": peaks.get("hbm_gbs"), "note": "HBM is not the bound (SURVEY.md 8d)"}}
        if front_name == "ddc_front_tc_kernel":
            # The tensor-core kernel is bound by the shared-memory pipe: one NCO table look-up per channel-sample, 32 lanes at 32
            # unrelated addresses.  A gather of 32 uniformly random words needs as many wavefronts as its fullest bank holds
            # words: E[max load of 32 balls in 32 bins], a property of the access pattern and not of the code.
            rng = np.random.default_rng(1)
            e_max = float(np.mean([np.bincount(rng.integers(0, 32, 32), minlength=32).max() for _ in range(20000)]))
            alg = units * e_max / 32.0 / front_s if front_s > 0 else 0.0
            executed = (static["lsu_wavefronts_per_unit"] * units / front_s) if static.get("lsu_wavefronts_per_unit") and front_s > 0 else None
            roofline = dict(common, **{
                "bound": "smem", "achieved": alg / 1e9, "peak": lds_peak / 1e9, "unit": "G wavefronts/s (shared-memory pipe)",
                "frac": (alg / lds_peak) if lds_peak else None,
                "algorithmic_wavefronts_per_unit": e_max / 32.0, "expected_max_bank_load": e_max,
                "executed_wavefronts_per_unit": static.get("lsu_wavefronts_per_unit"),
                "pipe_busy_frac": (executed / lds_peak) if executed and lds_peak else None,
                "pipe_busy_ncu_pct": static.get("lsu_pipe_pct_ncu"),
                "tensor": {"mma_per_unit": 4.0 / (32 * 128), "shape": "M128 N16 K32 kind::i8", "tensor_pipe_pct_ncu": static.get("tensor_pipe_pct_ncu"),
                           "note": "the integrators are an exact int8 GEMM (ddc_front_tc.cuh); the tensor pipe idles - the NCO look-ups bound the kernel"},
                "int32_equivalent": {"ops_per_unit": A_INT_OPS, "top_s": achieved / 1e12, "int32_peak_top_s": int32_peak / 1e12,
                                     "note": "the 41 INT32 ops per channel-sample of the HDL-literal formulation (SURVEY.md 8d) against the "
                                             "CUDA-core peak: above 1 because 30 of them (the integrator adds) now run as int8 MMAs"},
                "note": "achieved = channel-samples x E[max bank load]/32 wavefronts (the conflict-limited gather of one table word per "
                        "channel-sample) / kernel time; pipe_busy_frac adds the ADC broadcasts and the record stores "
                        "(executed_wavefronts_per_unit, from ncu); issue_frac = executed warp instructions x 32 / CUDA-core peak",
                "peak_source": "ua3reo_measure_lds_peak (conflict-free LDS.32 streams, 1024 threads per SM), measured live on this GPU; "
                               "MEASURED_PEAKS.json has no shared-memory peak"})
        else:
            roofline = dict(common, **{
                "bound": "int32_alu", "achieved": achieved / 1e12, "peak": int32_peak / 1e12,
                "unit": "TOP/s (INT32)", "frac": (achieved / int32_peak) if int32_peak else None,
                "note": "achieved counts the 41 algorithmic INT32 ops per channel-sample of the HDL-literal formulation "
                        "(SURVEY.md 8d), which is why frac can exceed 1: the kernel executes executed_instructions_per_unit "
                        "of them (table-driven NCO output stage, shifts fused into adds); issue_frac = executed ops / peak "
                        "is the fraction to judge the kernel by",
                "ops_per_unit": A_INT_OPS,
                "peak_source": "ua3reo_measure_int32_peak (IMAD+LOP3+IADD3 chains), measured live on this GPU; "
                               "MEASURED_PEAKS.json has no integer peak"})
        line = {
            "metric": "rx_chain_channel_adc_samples_per_s" if full else "ddc_channel_adc_samples_per_s", "value": value,
            "unit": "channel*samples/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "int64+f32" if full else "int64", "data": "synthetic",
            "config": {"workload": workload_s,
                       "channels_total": n_ch * world, "channels_per_gpu": n_ch, "block_samples": block,
                       "real_time_channels": value / 49152000.0,
                       "clocking_class": "B/3/129 (the board's most frequent frame alignment, DESIGN.md 2)",
                       "comm_sms": comm_sms, "adc_transport": transport, "spectra_gather": gather_transport,
                       "l2": "per-step working set ~%.0f MB (chunk records + frames) exceeds the 126 MB L2; no flush needed"
                             % (step_bytes / 1e6)},
            "e2e": {"value": e2e, "unit": "channel*samples/s", "h2d_bytes_per_step": 2 * block,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / K, "host_enqueue_ms_per_step": host_enqueue_ms[0],
                    "host_d2h_aggregate_gbs": world * 4.0 * (64 << 20) / (red[4] * 1e-3) / 1e9 if red[4] > 0 else None,
                    "d2h_gbs_needed_at_device_rate": d2h * world / (ms / K * 1e-3) / 1e9,
                    "note": "bytes are per rank; host_d2h_aggregate_gbs = all ranks copying device -> pinned host at once "
                            "(4 x 64 MB each), the box's ceiling for results leaving the GPUs"},
            "gpu_launches": int(launches),
            "parity": {"ddc_ranks_ok": int(oks[0]), "ranks": world, "channels_checked_per_rank": len(picks),
                       "ddc_check": "frames of one more block (received from rank 0 at N>1) == golden model, bit for bit",
                       "stm32_ranks_ok": (int(oks[1]) if full and audio_ok is not None else None), "stm32_check": audio_note,
                       "spectra_gather_ok": gather_ok,
                       "oracle": "golden C model pinned edge for edge to the reference's executed VHDL (filters) and Verilog (mixers, "
                                 "shifts, Q delay, DAC corrector, MCU bus); the NCO's output rounding / start phase is a documented "
                                 "convention (its Altera submodules are encrypted), so LSB parity with silicon is not claimed for it"},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        if sustained:
            line["sustained"] = {"value": float(n_ch) * world * block * sustained[1] / (red[2] * 1e-3), "unit": "channel*samples/s",
                                 "seconds": red[2] * 1e-3, "steps": sustained[1], "ms_per_step": red[2] / sustained[1],
                                 "clocks": sustained[2]}
        if tx_duc:
            dac = float(n_ch) * world * tx_duc[2] * 1024 * tx_duc[1]
            line["tx_duc"] = {"metric": "duc_channel_dac_samples_per_s", "value": dac / (red[3] * 1e-3), "unit": "channel*samples/s",
                              "ms_per_push": red[3] / tx_duc[1], "tx_samples_per_push": tx_duc[2], "gpu_launches": int(tx_duc[3]),
                              "real_time_channels": dac / (red[3] * 1e-3) / 49152000.0,
                              "note": "tx_ciccomp x2 + tx_cic x512 + tx_mixer + tx_summator + DAC_corrector for every channel "
                                      "(ua3reo_duc_push, host I/Q in); the board is half duplex, so this is timed on its own"}
    if (world > 1 and oks[0] != world) or (world == 1 and not ddc_ok):
        raise SystemExit("bench.py: PARITY FAILURE - CUDA frames differ from the golden model on %d of %d ranks" % (world - int(oks[0]), world))
    if gather_ok is False:
        raise SystemExit("bench.py: PARITY FAILURE - the spectra gathered on rank 0 differ from what the ranks computed")
    if full and audio_ok is not None and int(oks[1]) != world:
        raise SystemExit("bench.py: PARITY FAILURE - STM32-stage results outside tolerance (%s)" % audio_note)
    # tensors that were used on the library's streams must be released while those streams still exist
    if gather_spectra is not None:
        import gc
        torch.cuda.synchronize()
        if slab_gather is not None:
            slab_gather.close()                        # collective
        del gather_spectra, gather_spectra_ce, gather_spectra_nccl, slab_gather, spec_dev, spec_all, spec_all_host, cstream, gstream, ev_spec, ev_gdone, gather_pg
        gc.collect()
    torch.cuda.synchronize()
    if bc is not None and hasattr(bc, "close"):
        bc.close()                                     # collective: peers unmapped, barrier, arenas freed
    del bc
    rx.close()
    return line


def run_ours(args):
    import torch
    import torch.distributed as dist
    import ua3reo_loader
    pkg = ua3reo_loader.load()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - this repo has no CPU fallback (use --impl reference for the CPU model)")
    torch.cuda.set_device(local)
    if world > 1:
        # the broadcast moves 2 MB per step: one NCCL channel (= one CTA) is plenty, and that is how many SMs are kept for it
        os.environ.setdefault("NCCL_MAX_NCHANNELS", str(max(1, args.comm_sms if args.comm_sms is not None else 1)))
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    want_cpu = world == 1 and not args.no_cpu_baseline
    if args.scaling == "strong":
        assert args.total_channels % world == 0
        ddc_ch = args.total_channels // world
        full_ch = max(32, args.full_channels_per_gpu // world)      # configs[4] names 4096 channels in total
    else:
        ddc_ch = args.channels_per_gpu or CH_PER_GPU
        full_ch = args.full_channels_per_gpu
    line = None
    if args.workload in ("all", "ddc"):
        line = measure(args, pkg, torch, dist if world > 1 else None, "ddc", ddc_ch, world, rank, local, want_cpu)
    if args.workload in ("all", "full_chain"):
        fc = measure(args, pkg, torch, dist if world > 1 else None, "full_chain",
                     full_ch if args.workload == "all" or args.channels_per_gpu is None else args.channels_per_gpu,
                     world, rank, local, want_cpu)
        if rank == 0:
            if line is None:
                line = fc
            else:
                line["full_chain"] = fc
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def emit(line):
    """the ONE line on stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse()
    # stdout carries the JSON line and nothing else: libraries that print there (NCCL's version banner when the box sets
    # NCCL_DEBUG=VERSION, a stray warning) are sent to stderr at the file-descriptor level for the whole run
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
