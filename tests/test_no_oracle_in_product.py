"""The CPU oracle is test infrastructure: nothing shipped (st_ito_b200/, include/, scripts/) may import, load, link or
execute anything under oracle/; bench.py may only do so inside its CPU legs."""
import ast
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py_files(*dirs):
    for d in dirs:
        for base, _dirs, files in os.walk(os.path.join(ROOT, d)):
            for f in files:
                if f.endswith(".py"):
                    yield os.path.join(base, f)


def _imports(path):
    tree = ast.parse(open(path).read(), filename=path)
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                yield a.name, node.lineno
        elif isinstance(node, ast.ImportFrom) and node.module:
            yield node.module, node.lineno


def test_product_python_never_imports_the_oracle():
    offenders = []
    for path in list(_py_files("st_ito_b200", "scripts")):
        for mod, line in _imports(path):
            if mod == "oracle" or mod.startswith("oracle."):
                offenders.append(f"{os.path.relpath(path, ROOT)}:{line}")
        src = open(path).read()
        if re.search(r"liboracle|oracle/_build|dsp_oracle", src):
            offenders.append(os.path.relpath(path, ROOT) + " (mentions the oracle library)")
    assert not offenders, offenders


def test_native_sources_and_library_do_not_touch_the_oracle():
    for base, _dirs, files in os.walk(os.path.join(ROOT, "st_ito_b200", "csrc")):
        for f in files:
            if f.endswith((".cu", ".h", ".cuh", ".cpp")) or f == "Makefile":
                src = open(os.path.join(base, f)).read()
                assert "#include \"../../oracle" not in src and "liboracle" not in src, f
    lib = os.path.join(ROOT, "st_ito_b200", "libstito.so")
    if os.path.isfile(lib):
        needed = subprocess.check_output(["readelf", "-d", lib], text=True)
        assert "oracle" not in needed
        syms = subprocess.check_output(["nm", "-D", lib], text=True)
        assert "oracle_" not in syms


def test_bench_imports_the_oracle_only_inside_its_cpu_legs():
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    allowed = {"cpu_candidates", "cpu_setup"}
    for node in tree.body:
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            mods = [a.name for a in node.names] if isinstance(node, ast.Import) else [node.module or ""]
            assert not any(m == "oracle" or m.startswith("oracle.") for m in mods), "module-level oracle import in bench.py"
        if isinstance(node, ast.FunctionDef):
            uses = any((isinstance(n, ast.ImportFrom) and (n.module or "").split(".")[0] == "oracle") or
                       (isinstance(n, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in n.names))
                       for n in ast.walk(node))
            assert not uses or node.name in allowed, f"bench.py:{node.name} imports the oracle"
