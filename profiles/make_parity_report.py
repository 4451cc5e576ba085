#!/usr/bin/env python
"""profiles/r02_parity.json from the JSONL the GPU tests write (tests/parity.py::log_record, $DDP_PARITY_LOG).

    python profiles/make_parity_report.py gpurun_out/<run>/parity_log.jsonl profiles/r02_parity.json "<commit / command>"

One entry per parity check that ran on the B200: the rule that decided it ("exact" = every logit within ATOL and every
class-map pixel identical to the fp32 oracle run open-loop; "closed_loop" = adjudicated step by step with the fp64
oracle), mismatch counts, max |delta| and the fp32-vs-fp64 floors, so the evidence survives `pytest -q`."""
import json
import sys


def main(src, dst, note):
    recs = [json.loads(l) for l in open(src) if l.strip()]
    summary = {
        "source": note,
        "records": len(recs),
        "exact": sum(r["rule"] == "exact" for r in recs),
        "closed_loop": sum(r["rule"] == "closed_loop" for r in recs),
        "closed_loop_with_flips": sum(r["rule"] == "closed_loop" and r.get("flips", 0) > 0 for r in recs),
        "total_flipped_pixel_steps": sum(r.get("flips", 0) for r in recs if r["rule"] == "closed_loop"),
        "total_pixel_steps_closed_loop": sum(r.get("pixel_steps", 0) for r in recs if r["rule"] == "closed_loop"),
        "max_abs_d_out_exact": max([r.get("max_abs_d_out", 0.0) for r in recs if r["rule"] == "exact"] or [0.0]),
        "max_abs_d_logits_closed_loop": max([r.get("max_abs_d_logits", 0.0) for r in recs if r["rule"] == "closed_loop"] or [0.0]),
        "max_flip_margin": max([r.get("max_flip_margin", 0.0) for r in recs] or [0.0]),
        "baseline_shapes": [r for r in recs if "BASELINE" in r.get("what", "")],
    }
    json.dump({"summary": summary, "records": recs}, open(dst, "w"), indent=1)
    print(json.dumps({k: v for k, v in summary.items() if k != "baseline_shapes"}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
