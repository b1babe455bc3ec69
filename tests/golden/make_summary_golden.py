"""Golden vectors for bamqc_b200/summary.py: run the REFERENCE post-processor (/root/reference/bamqc_summary.py,
imported as a module, unmodified) on the committed .bamqc fixtures and record every summary field, the three flag
sets, the dense line and the long text (or the exception the reference raises).  Run in the build container only;
the GPU box has no /root/reference.  Usage: python tests/golden/make_summary_golden.py"""
import contextlib
import importlib.util
import io
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_bamqc_summary", "/root/reference/bamqc_summary.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def capture(fn, *a):
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            fn(*a)
        return buf.getvalue(), None
    except Exception as e:  # the reference's own failure modes are part of the record
        return buf.getvalue(), type(e).__name__


out = {}
for name in sorted(os.listdir(HERE)):
    if not name.endswith(".bamqc"):
        continue
    lanes = []
    data = []
    try:
        ref.read_bamqc_output(data, os.path.join(HERE, name))
    except Exception as e:
        out[name] = {"read_error": type(e).__name__}
        continue
    for lane in data:
        summary = {}
        try:
            ref.summarize(summary, lane)
        except Exception as e:
            lanes.append({"summarize_error": type(e).__name__})
            continue
        rec = {"fields": {k: v for k, v in summary.items() if not isinstance(v, (list, dict, set))}}
        for st in (3, 2, 1):
            summary["flags%d" % st] = ref.get_flags(summary, st)
            rec["flags%d" % st] = sorted(summary["flags%d" % st])
        rec["dense"], rec["dense_error"] = capture(ref.write_line, summary)
        rec["long"], rec["long_error"] = capture(ref.write_txt, summary)
        lanes.append(rec)
    out[name] = {"lanes": lanes}
hdr, _ = capture(ref.write_header)
out["__header__"] = hdr
json.dump(out, open(os.path.join(HERE, "summary", "summary_golden.json"), "w"), indent=1, sort_keys=True)
print("wrote", len(out) - 1, "files")
