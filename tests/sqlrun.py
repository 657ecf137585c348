"""Run SQL statements against an extension library in a SEPARATE process (one library per process:
SQLite dlopens extensions RTLD_GLOBAL and the reference exports every symbol, SURVEY trap 2)."""
import json
import os
import subprocess
import sys

CHILD = r'''
import json, sqlite3, sys
lib, stmts = sys.argv[1], json.loads(sys.stdin.read())
con = sqlite3.connect(":memory:", isolation_level=None)
con.enable_load_extension(True)
con.load_extension(lib)
out = []
def enc(v):
    if isinstance(v, bytes): return {"hex": v.hex()}
    if isinstance(v, float):
        if v != v: return {"f": "nan"}
        if v in (float("inf"), float("-inf")): return {"f": str(v)}
    return v
for s in stmts:
    try:
        if isinstance(s, list):
            sql, params = s[0], [bytes.fromhex(p["hex"]) if isinstance(p, dict) else p for p in s[1]]
        else:
            sql, params = s, []
        rows = con.execute(sql, params).fetchall()
        out.append({"rows": [[enc(v) for v in r] for r in rows]})
    except Exception as e:
        out.append({"error": str(e)})
print("\n@@JSON@@" + json.dumps(out), flush=True)
'''


def run_sql(lib_path: str, statements, timeout=600):
    """lib_path without the .so suffix is fine (sqlite appends it).  Returns a list of {"rows": ...} | {"error": ...}."""
    p = subprocess.run([sys.executable, "-c", CHILD, lib_path], input=json.dumps(statements), capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"sql child failed: {p.stderr[-2000:]}")
    for ln in p.stdout.splitlines():          # the reference printf()s diagnostics to stdout; find our marker
        if ln.startswith("@@JSON@@"):
            return json.loads(ln[len("@@JSON@@"):])
    raise RuntimeError("no result line from sql child: " + p.stdout[-500:])


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "sqlite_vector_b200", "lib", "vector")
REF_CPU = os.path.join(ROOT, "oracle", "_ref", "cpu", "vector")


def blob(a):
    return {"hex": a.tobytes().hex()}
