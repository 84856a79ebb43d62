"""bench.py's host-side pieces, without a GPU: the workload tables, the contig length recipes and the roofline constants."""
import importlib.util
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_workloads_and_length_recipes():
    bench = _bench()
    for name, w in bench.WORKLOADS.items():
        lens = bench.contig_lengths(w)
        if w["shape"] == "conifer":
            assert abs(sum(lens) - w["total"]) < w["total"] // 100, name  # (the 200-base floor adds a little)
        else:
            assert sum(lens) == w["total"], name
        assert min(lens) >= 200, name
        assert bench.algorithmic_bytes_per_base(w) == (3553 if w.get("snv") and w["k"] == 32 else 97), name
    for name, stages in bench.NCU_TRAFFIC.items():
        assert name in bench.WORKLOADS
        for stage, (nbytes, source) in stages.items():
            assert stage in ("scan", "presite", "walk") and nbytes > 0
            assert os.path.exists(os.path.join(ROOT, source.split(":")[0])), source


def test_every_name_bench_uses_is_defined():
    """A cut that removes a helper shows up here, not on the GPU box (pyflakes-style check with the compiler's symbol table)."""
    import symtable
    src = open(os.path.join(ROOT, "bench.py")).read()
    top = symtable.symtable(src, "bench.py", "exec")
    defined = {s.get_name() for s in top.get_symbols() if s.is_assigned() or s.is_imported() or s.is_namespace()}
    import builtins
    missing = set()

    def walk(t):
        for s in t.get_symbols():
            if s.is_global() and s.is_referenced() and s.get_name() not in defined and not hasattr(builtins, s.get_name()) and s.get_name() not in ("__file__", "__name__"):
                missing.add(s.get_name())
        for c in t.get_children():
            walk(c)

    walk(top)
    assert not missing, sorted(missing)


def test_help_runs():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "--impl" in out.stdout
