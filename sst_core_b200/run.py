"""`sst`-like front end: python -m sst_core_b200 [options] model.py

Mirrors the reference command line for the options the hot path needs (config.cc:412-504): --model-options,
--stop-at, --num-threads/-n (accepted; partitions are GPUs here), --partitioner (only sst.linear), --timebase,
--lib-path (accepted, ignored: device elements are compiled into libsstb200.so), --output-directory.
"""
from __future__ import annotations

import argparse
import os
import sys
from typing import Optional

import numpy as np

from . import checkpoint, config, output, profile, wireup
from .elements import NRESULT, RESULT_WORDS, RESULT_WORDS_NOSTATS


def result_words(m) -> int:
    """How many words of the 32-word result record the model's elements fill."""
    table = RESULT_WORDS if m.stats_enabled else RESULT_WORDS_NOSTATS
    return max((table.get(int(t), NRESULT) for t in _types_present(m)), default=NRESULT)


def _types_present(m):
    t = m.comp_type
    if t.size and bool((t == t[0]).all()):  # the usual large model: one element type (np.unique would sort 16 M ids)
        return [t[0]]
    return np.unique(t)


def simulate_model(m: wireup.WiredModel, device: int = 0, auto_tune: bool = True, extras: Optional[dict] = None,
                   checkpoint_period: int = 0, checkpoint_writer=None, restore_blob=None, **engine_opts):
    """Run a WiredModel on one GPU.  Returns (results [ncomp,32] u64, Status).  With ``handler_counts=True`` and an
    ``extras`` dict, ``extras["handler_counts"]`` receives (deliveries per port, clock calls per component).
    ``checkpoint_period`` (cycles) + ``checkpoint_writer`` (checkpoint.CheckpointWriter): write a checkpoint every
    period of simulated time; ``restore_blob``: continue from a checkpoint instead of running setup().

    The calendar geometry (window slots, bin and staging capacities) is a tuning choice; when a run reports that one of
    them overflowed and ``auto_tune`` is set, the run is repeated from the start with that resource doubled (an
    overflow is always detected and reported, never silent -- see Engine.raise_on_error)."""
    from .engine import Engine, EngineError

    opts = dict(engine_opts)
    for attempt in range(10):
        e = Engine(m, device=device, **opts)
        try:
            if restore_blob is not None:
                e.restore(restore_blob)
            else:
                e.setup()
            try:
                if m.stat_rate and m.stats_enabled and extras is not None:
                    # periodic statistic output: the first at `rate`, then every `rate` (a statistic clock)
                    extras["periodic"] = []
                    t_next = (e.status().now // m.stat_rate + 1) * m.stat_rate
                    while True:
                        rec, fin = e.run_to_report(t_next, result_words(m))
                        if rec is None:
                            break
                        extras["periodic"].append((t_next, rec))
                        t_next += m.stat_rate
                        if fin:
                            break
                    if checkpoint_period:
                        raise EngineError("periodic statistic output and checkpoints cannot be combined yet")
                    e.run()
                elif checkpoint_period:
                    if checkpoint_writer is not None:
                        checkpoint_writer.next_id = 1 if restore_blob is None else checkpoint_writer.next_id
                    checkpoint.run_with_checkpoints(e, checkpoint_period, checkpoint_writer, opts)
                else:
                    e.run()
            except EngineError:
                if not e.status().error:
                    raise
            st = e.status()
            if st.error in (1, 2) and auto_tune and attempt < 9 and restore_blob is None:
                if st.error == 1:  # calendar bin overflow: more slots first (events spread over more bins), then bigger bins
                    if opts.get("calendar_slots", 2) < 16:
                        opts["calendar_slots"] = 2 * opts.get("calendar_slots", 2)
                    else:
                        opts["bin_capacity"] = 2 * (opts.get("bin_capacity", 0) or 2048)
                else:  # shared-memory staging overflow
                    opts["smem_events"] = min(2 * (opts.get("smem_events", 0) or 1280), 6144)
                    opts["bin_capacity"] = max(opts.get("bin_capacity", 0) or 2048, 2 * opts["smem_events"])
                continue
            e.raise_on_error(st)
            if extras is not None and opts.get("handler_counts"):
                extras["handler_counts"] = e.handler_counts()
                extras["sync_profile"] = [e.sync_profile()]
                if int(opts["handler_counts"]) & 2:
                    extras["handler_times"] = e.handler_times()
            if extras is not None and ((m.output_fmt and any(f is not None for f in m.output_fmt)) or m.stat_slots):
                extras["output"] = e.output_records()
            return e.results(result_words(m)), st
        finally:
            e.close()
    raise EngineError("auto-tuning did not converge")


def simulate_distributed(m: wireup.WiredModel, extras: Optional[dict] = None, checkpoint_period: int = 0,
                         checkpoint_writer=None, restore_registry: Optional[dict] = None, **engine_opts):
    """Run a WiredModel partitioned over the ranks of the initialised torch.distributed world (one GPU per rank).
    Returns (results of THIS rank's components, Status, partitioned model).  ``extras["handler_counts"]`` (with
    ``handler_counts=True``): this rank's counters.  ``checkpoint_period`` / ``checkpoint_writer(rank, n, pm)``: every
    rank writes its partition file, rank 0 the registry.  ``restore_registry``: continue from a checkpoint -- taken on
    any number of ranks -- instead of running setup()."""
    import torch
    import torch.distributed as dist

    from .engine import DistributedRunner, Engine, EngineError, connect_peers, partition_model

    n, r = dist.get_world_size(), dist.get_rank()
    pm = partition_model(m, n)
    dev = torch.cuda.current_device()
    # the engine's kernels and the NCCL collectives must be ordered on ONE stream: torch orders its collectives after
    # the *current* stream, so make a dedicated stream current and hand the same stream to the engine
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        e = Engine(pm, device=dev, rank=r, n_ranks=n, stream=stream.cuda_stream, **engine_opts)
        try:
            connect_peers(e)  # NVLink peer mode when the job allows it (collective; before setup / restore)
            if restore_registry is not None:
                e.restore(checkpoint.blob_for_rank(restore_registry, pm, r, n, engine_opts.get("outbox_capacity", 0)))
            else:
                e.setup()
            runner = DistributedRunner(e)
            if pm.stat_rate and pm.stats_enabled and extras is not None:
                if checkpoint_period:
                    raise EngineError("periodic statistic output and checkpoints cannot be combined yet")
                extras["periodic"] = []
                t_next = (e.status().now // pm.stat_rate + 1) * pm.stat_rate
                while True:
                    rec, fin = runner.run_to_report(t_next, result_words(pm))
                    if rec is None:
                        break
                    extras["periodic"].append((t_next, rec))
                    t_next += pm.stat_rate
                    if fin:
                        break
                runner.run()
            elif checkpoint_period:
                writer = checkpoint_writer(r, n, pm) if checkpoint_writer else None
                checkpoint.run_with_checkpoints(e, checkpoint_period, writer, engine_opts, runner=runner)
            else:
                runner.run()
            st = e.status()
            e.raise_on_error(st)
            if extras is not None and engine_opts.get("handler_counts"):
                extras["handler_counts"] = e.handler_counts()
                extras["sync_profile"] = [e.sync_profile()]
                if int(engine_opts["handler_counts"]) & 2:
                    extras["handler_times"] = e.handler_times()
            return e.results(result_words(pm)), st, pm
        finally:
            e.close()


def main(argv: Optional[list] = None) -> int:
    ap = argparse.ArgumentParser(prog="sst_core_b200", description="B200 engine behind SST-core's model surface")
    ap.add_argument("script")
    ap.add_argument("--model-options", default="")
    ap.add_argument("--stop-at", default=None)
    ap.add_argument("-n", "--num-threads", type=int, default=1)
    ap.add_argument("--partitioner", default="sst.linear")
    ap.add_argument("--timebase", default=None)
    ap.add_argument("--lib-path", default=None)
    ap.add_argument("--output-directory", default=".")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--calendar-slots", type=int, default=0)
    ap.add_argument("--bin-capacity", type=int, default=0)
    ap.add_argument("--checkpoint-sim-period", "--checkpoint-period", default="", metavar="PERIOD",
                    help="write a checkpoint every PERIOD of simulated time (e.g. 100ns)")
    ap.add_argument("--checkpoint-prefix", default="checkpoint")
    ap.add_argument("--load-checkpoint", action="store_true",
                    help="the input file is a checkpoint registry (*.sstcpt): restart from it")
    ap.add_argument("--output-json", default="", metavar="FILE",
                    help="write the configuration graph in the reference's JSON interchange format (sst model.json)")
    ap.add_argument("--run-mode", default="both", choices=["init", "run", "both"],
                    help="init: build (and write) the graph only")
    ap.add_argument("--enable-profiling", default="",
                    help="name:type(key=value,...)[point,...];...  (handler count tools, see profile.py)")
    ap.add_argument("--profiling-output", default="", help="file for the profiling records (default: stdout)")
    a = ap.parse_args(argv)
    if a.partitioner not in ("sst.linear", "linear"):
        print(f"FATAL: partitioner {a.partitioner} is not supported (only sst.linear)", file=sys.stderr)
        return 1
    reg = None
    try:
        if a.load_checkpoint:
            # simulation.cc:2073-2200: the run's configuration comes from the checkpoint, not from the command line
            reg = checkpoint.load_registry(a.script)
            rec = reg["recipe"]
            a.script, a.model_options, a.stop_at, a.timebase = (rec["script"], rec["model_options"], rec["stop_at"],
                                                                rec["timebase"])
            a.enable_profiling = a.enable_profiling or rec.get("enable_profiling", "")
        if a.script.endswith(".json"):
            g = config.load_json(a.script)
        else:
            g = config.build_graph(a.script, a.model_options)
        if a.timebase:
            g.program_options["timebase"] = a.timebase
        if a.output_json:
            if a.stop_at:
                g.program_options["stop-at"] = a.stop_at
            config.dump_json(g, os.path.join(a.output_directory, a.output_json))
        if a.run_mode == "init":
            return 0
        m = wireup.wire(g, stop_at_override=a.stop_at)
        if reg is not None and checkpoint.model_fingerprint(m) != reg["model_fingerprint"]:
            raise config.ModelError(f"the model rebuilt from {a.script} is not the one the checkpoint was taken from")
        tools = profile.parse_enable_profiling(a.enable_profiling) if a.enable_profiling else []
    except config.ModelError as ex:
        print(f"FATAL: {ex}", file=sys.stderr)
        return 1
    opts = {}
    if a.calendar_slots:
        opts["calendar_slots"] = a.calendar_slots
    if a.bin_capacity:
        opts["bin_capacity"] = a.bin_capacity
    extras = {}
    if tools:
        opts["handler_counts"] = 3 if profile.needs_times(tools) else 1
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ckpt = {}
    recipe = None
    if reg is not None:
        opts = dict(reg["engine_opts"])
    if a.checkpoint_sim_period:
        from .timelord import TimeLord
        recipe = {"script": os.path.abspath(a.script), "model_options": a.model_options, "stop_at": a.stop_at,
                  "timebase": a.timebase, "enable_profiling": a.enable_profiling}
        ckpt["checkpoint_period"] = TimeLord(m.timebase).cycles(a.checkpoint_sim_period)

    wired = m  # the fingerprint in the registry is that of the model as wired from the script, before any renumbering

    def make_writer(rank, n_ranks, _partitioned=None):
        w = checkpoint.CheckpointWriter(a.output_directory, a.checkpoint_prefix, recipe, wired, rank=rank, n_ranks=n_ranks)
        if reg is not None:
            w.next_id = reg["checkpoint_id"] + 1
        return w

    if world > 1:
        # under torchrun: one process per GPU, the model cut by sst.linear, results gathered on rank 0
        #   python -m torch.distributed.run --nproc-per-node N -m sst_core_b200 [options] model.py
        import torch
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        try:
            if recipe is not None:
                ckpt["checkpoint_writer"] = make_writer
            # m: renumbered so that every rank owns a contiguous range
            part, st, m = simulate_distributed(m, extras=extras, restore_registry=reg, **ckpt, **opts)
            parts = [None] * world if dist.get_rank() == 0 else None
            dist.gather_object((part, extras.get("handler_counts"), extras.get("periodic", []),
                                extras.get("handler_times"), extras.get("sync_profile")), parts, dst=0)
            if dist.get_rank() != 0:
                return 0
            res = np.concatenate([p[0] for p in parts], axis=0)
            extras["periodic"] = [(t, np.concatenate([p[2][i][1] for p in parts], axis=0))
                                  for i, (t, _) in enumerate(parts[0][2])]
            if tools:
                extras["handler_counts"] = tuple(np.concatenate([p[1][i] for p in parts]) for i in (0, 1))
                extras["sync_profile"] = [p[4][0] for p in parts]
                if parts[0][3] is not None:
                    extras["handler_times"] = tuple(np.concatenate([p[3][i] for p in parts]) for i in (0, 1))
        finally:
            dist.barrier()
            dist.destroy_process_group()
    else:
        if reg is not None:
            ckpt["restore_blob"] = checkpoint.blob_for_rank(reg, m, 0, 1)
        if recipe is not None:
            ckpt["checkpoint_writer"] = make_writer(0, 1, m)
        res, st = simulate_model(m, device=a.device, extras=extras, **ckpt, **opts)
    for lines in (m.construct_console or []):  # what the constructors print (before anything runs)
        for line in lines:
            print(line)
    for line in output.console_lines(m, extras.get("output"), st.end_time):
        print(line)
    if m.stat_slots and extras.get("output") is not None:
        # statistics under the device statistics engine: periodic / event-triggered outputs from the log, then the
        # end-of-simulation output of every statistic
        rec = np.concatenate([extras["output"], output.stat_end_records(m, res, st.end_time)])
        console, files = output.stat_engine_outputs(m, rec, st.end_time)
        for line in console:
            print(line)
        for path, lines in files.items():
            with open(os.path.join(a.output_directory, path), "w") as f:
                f.write("\n".join(lines) + "\n")
    for line in output.finish_lines(m, res, st.end_time):
        print(line)
    if tools:
        recv, calls = extras["handler_counts"]
        text = profile.profiling_text(m, tools, recv, calls, st.end_time, os.path.abspath(a.script),
                                      times=extras.get("handler_times"), sync=extras.get("sync_profile"))
        if a.profiling_output:
            with open(os.path.join(a.output_directory, a.profiling_output), "w") as f:
                f.write(text)
        else:
            sys.stdout.write(text)
    if m.stat_output[0] == "sst.statOutputConsole":
        for line in output.stat_lines_console(m, res):
            print(line)
    rows = output.stat_rows_csv(m, res, st.end_time, periodic=extras.get("periodic", ()))
    if rows and m.stat_output[0] == "sst.statOutputCSV":
        path = os.path.join(a.output_directory, m.stat_output[1].get("filepath", "StatisticOutput.csv"))
        with open(path, "w") as f:
            f.write("\n".join(rows) + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
