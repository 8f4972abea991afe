"""GPU timeline of an executed program in the schema of the reference simulator's
``chrome_timeline.json`` (vTCAD/code/simulator.py:360-382: a JSON list of ``ph: "X"`` events with
``name`` = instruction TYPE, ``cat`` = instruction ID, ``ts``/``dur``, ``pid`` = ``tid`` = hardware
unit), so a modelled ASIC timeline and a measured B200 timeline load side by side in chrome://tracing.

    kernels.EVENT_LOG = []
    out, log = executor.execute(..., return_log=True)
    trace.save_timeline(kernels.EVENT_LOG, log, "Results/GPU")
"""
from __future__ import annotations

import json
import os

import torch

#: kernel entry point -> (instruction TYPE as the ISA names it, hardware unit the simulator would bind)
_UNIT = {
    "gta_gemm_f32": ("COMP_MM", "MM"),
    "gta_aggregate_f32": ("COMP_MUL_COMP_ADD", "VEC_ALU"),
    "gta_gat_aggregate_f32": ("COMP_ADD_COMP_SF_COMP_MUL_COMP_ADD", "VEC_ALU"),
    "gta_gat_logits_f32": ("COMP_ADD_COMP_SF", "SF_ALU"),
    "nccl_all_gather": ("LOAD_N", "Memory_Access_Unit"),
}


def timeline_events(event_log, kernel_log=None, unit_scale: float = 1e3) -> list:
    """``event_log``: kernels.EVENT_LOG entries ``(name, start_event, end_event)``; ``kernel_log``: the
    ``(kernel, op position)`` list of ``execute(return_log=True)`` (gives the ``cat`` field an op id).
    Times are microseconds from the first event (``ts``, ``dur`` as in the simulator, where the unit
    is one cycle = 1 ns)."""
    if not event_log:
        return []
    torch.cuda.synchronize()
    t0 = event_log[0][1]
    ops = {}
    for kernel, pos in (kernel_log or []):
        ops.setdefault(kernel.split(":")[0].split("+")[0], []).append(pos)
    seen = {}
    events = []
    for name, start, end in event_log:
        typ, unit = _UNIT.get(name, (name, "VEC_ALU"))
        k = seen.get(name, 0)
        seen[name] = k + 1
        pos = ops.get(name, [])
        cat = f"{pos[k % len(pos)]}_{name}" if pos else name
        events.append({"name": typ, "cat": cat, "ph": "X", "ts": t0.elapsed_time(start) * unit_scale,
                       "dur": start.elapsed_time(end) * unit_scale, "pid": unit, "tid": unit})
    return events


def save_timeline(event_log, kernel_log, folder_name: str) -> str:
    os.makedirs(folder_name, exist_ok=True)
    path = os.path.join(folder_name, "chrome_timeline.json")
    with open(path, "w") as f:
        json.dump(timeline_events(event_log, kernel_log), f, indent=4)
    return path
