"""bench.py contract checks that run without a GPU: the reference arm (the reference's CPU path timed on the host cores) prints ONE
JSON line with the keys the driver reads, and the torchrun convention (only rank 0 works) holds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--no-config1"],
                       capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout.strip()


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run()
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "denoising_steps_per_sec" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    # "reference" where the reference tree (or its staged copy oracle/_ref) is present, the oracle port otherwise
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["batch_per_gpu"] == 256       # the true batch of the GPU arm, no extrapolation
    assert abs(d["cpu_baseline"]["value"] - d["value"]) < 1e-9
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_without_work():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == ""


def test_gpu_arm_line_has_every_contract_key():
    """Static check (no GPU): the dict literal bench.py prints on rank 0 of the GPU arm carries every key of the bench
    contract, and every name its value expressions use is bound in that function (a typo would only surface on the box)."""
    import ast
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    found = None
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            if isinstance(node, ast.Assign) and isinstance(node.value, ast.Dict) and \
                    any(isinstance(k, ast.Constant) and k.value == "roofline_hbm" for k in node.value.keys):
                found = (fn, node.value)
    assert found, "the GPU arm's JSON dict was not found in bench.py"
    fn, d = found
    keys = {k.value for k in d.keys if isinstance(k, ast.Constant)}
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "roofline_hbm", "roofline_attention"}
    assert need <= keys, sorted(need - keys)
    sub = {k.value: v for k, v in zip(d.keys, d.values) if isinstance(k, ast.Constant)}
    for name in ("roofline", "roofline_attention"):
        rk = {k.value for k in sub[name].keys if isinstance(k, ast.Constant)}
        assert {"bound", "achieved", "peak", "unit", "frac"} <= rk, (name, rk)
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= {k.value for k in sub["e2e"].keys}
    # names read inside the dict must be assigned (or be parameters / imports / builtins) in the enclosing function or module
    bound = {a.arg for a in fn.args.args} | set(dir(__builtins__) if not isinstance(__builtins__, dict) else __builtins__)
    for n in ast.walk(fn):
        if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Store):
            bound.add(n.id)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            bound |= {(a.asname or a.name).split(".")[0] for a in n.names}
        elif isinstance(n, (ast.FunctionDef, ast.Lambda)) and n is not fn:
            bound |= {a.arg for a in n.args.args}
    for n in ast.iter_child_nodes(tree):
        if isinstance(n, (ast.FunctionDef, ast.ClassDef)):
            bound.add(n.name)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            bound |= {(a.asname or a.name).split(".")[0] for a in n.names}
        elif isinstance(n, ast.Assign):
            bound |= {t.id for t in n.targets if isinstance(t, ast.Name)}
    used = {n.id for n in ast.walk(d) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load)}
    assert used <= bound, sorted(used - bound)


def test_bench_describes_every_operand_scheme_it_can_be_switched_to():
    """REGEN_PRECISION selects the GPU arm's operand scheme; the `dtype` string of the line is looked up per scheme."""
    import re
    src = open(os.path.join(ROOT, "bench.py")).read()
    cm = open(os.path.join(ROOT, "regennet_b200", "cmdm.py")).read()
    schemes = set(re.findall(r"'(\w+)': \d", re.search(r"_PRECISIONS = \{([^}]*)\}", cm).group(1)))
    assert {"bf16x3", "mixed8", "mixed8h"} <= schemes
    dtypes = set(re.findall(r'^    "(\w+)": ', re.search(r"DTYPES = \{(.*?)\n\}", src, re.S).group(1), re.M))
    assert schemes - {"bf16"} <= dtypes, (schemes, dtypes)          # 'bf16' (single pass, ~1e-2) is outside the path's tolerance
    default = re.search(r'os\.environ\.get\("REGEN_PRECISION", "(\w+)"\)', src).group(1)
    assert default in dtypes
