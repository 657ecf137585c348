"""SQL scripts shared by the golden generator and the SQL-surface tests."""
import numpy as np

from oracle import pyoracle as po
from tests.sqlrun import blob


def surface_script():
    """no scan: encoders, option parsing, errors, quantization bytes and metadata"""
    rng = np.random.Generator(np.random.PCG64(2026))
    s = [
        "SELECT hex(vector_as_f32('[1.0, 2.5, -3e2, 0.1]'))",
        "SELECT hex(vector_as_f16('[1.0, 2.5, 65504, 1e-7, -0.0, 70000]'))",
        "SELECT hex(vector_as_bf16('[1.0, 2.5, 3.14159, -1e30]'))",
        "SELECT hex(vector_as_i8('[-1, 2, 127, -128, 3.9]'))",
        "SELECT hex(vector_as_u8('[1, 255, 0, 7.9]'))",
        "SELECT hex(vector_as_f32(' [ 1 , 2 , ] '))",
        "SELECT hex(vector_as_f32('[1,2,3]', 3))",
        "SELECT hex(vector_as_f32('[1,2,3]', 4))",
        "SELECT hex(vector_as_f32('1,2,3'))",
        "SELECT hex(vector_as_f32('[1,x]'))",
        "SELECT hex(vector_as_f32('[1 2]'))",
        "SELECT hex(vector_as_u8('[256]'))",
        "SELECT hex(vector_as_i8('[-129]'))",
        "SELECT hex(vector_as_f32(x'0000803F00000040'))",
        "SELECT hex(vector_as_f32(x'0000803F000000'))",
        "SELECT hex(vector_as_f32(x'0000803F00000040', 3))",
        "SELECT hex(vector_as_f16(x'003C0041', 2))",
        "SELECT vector_as_f32(42)",
        "SELECT vector_as_f32(NULL)",
        "SELECT vector_init('nope', 'e', 'type=FLOAT32,dimension=4')",
        "CREATE TABLE t (id INTEGER PRIMARY KEY, e BLOB, label TEXT, n INTEGER)",
        "SELECT vector_init('t', 'zzz', 'type=FLOAT32,dimension=4')",
        "SELECT vector_init('t', 'label', 'type=FLOAT32,dimension=4')",
        "SELECT vector_init('t', 'e', 'type=FLOAT99,dimension=4')",
        "SELECT vector_init('t', 'e', 'type=FLOAT32,dimension=-4')",
        "SELECT vector_init('t', 'e', 'type=FLOAT32')",
        "SELECT vector_init('t', 'e', 'type=FLOAT32,dimension=8,distance=nonsense')",
        "SELECT vector_init('t', 'e', 'type=FLOAT32,dimension=8,qtype=INT4')",
        "SELECT vector_init('t', 'e', 1)",
        "SELECT vector_init('t', 'e')",
        "SELECT vector_quantize('t', 'e')",
        "SELECT vector_quantize_preload('t', 'e')",
        "SELECT vector_init('t', 'e', ' type = float32 , dimension = 8, bogus=1, distance=COSINE,, =3 ')",
        "SELECT vector_init('t', 'e', 'type=FLOAT32,dimension=16')",
        "SELECT vector_init('t', 'e', 'type=FLOAT16,dimension=8')",
        "SELECT vector_init('t', 'e', 'type=FLOAT32,dimension=8,normalized=1')",
        "SELECT vector_init('t', 'e', 'type=FLOAT32,dimension=8')",
    ]
    x = rng.standard_normal((40, 8)).astype(np.float32)
    for i in range(40):
        if i == 7:
            s.append(["INSERT INTO t(id, e) VALUES (?, NULL)", [100 + 3 * i]])
        else:
            s.append(["INSERT INTO t(id, e) VALUES (?, ?)", [100 + 3 * i, blob(x[i])]])
    s += [
        "SELECT vector_quantize('t', 'e')",
        "SELECT rowid1, rowid2, counter, hex(data) FROM vector0_t_e",
        "SELECT tblname, colname, key, value FROM _sqliteai_vector ORDER BY key",
        "SELECT vector_quantize_memory('t', 'e')",
        "SELECT vector_quantize('t', 'e', 'max_memory=1KB')",
        "SELECT rowid1, rowid2, counter, length(data), hex(substr(data,1,24)) FROM vector0_t_e",
        "SELECT vector_quantize('t', 'e', 'qtype=UINT8,max_memory=0')",
        "SELECT rowid1, rowid2, counter, hex(data) FROM vector0_t_e",
        "SELECT tblname, colname, key, value FROM _sqliteai_vector ORDER BY key",
        "SELECT vector_quantize('t', 'e', 'qtype=INT8')",
        "SELECT key, value FROM _sqliteai_vector ORDER BY key",
        "SELECT vector_quantize_cleanup('t', 'e')",
        "SELECT count(*) FROM sqlite_master WHERE name='vector0_t_e'",
        "SELECT vector_quantize_memory('t', 'e')",
        # other source types
        "CREATE TABLE h (id INTEGER PRIMARY KEY, e BLOB)",
        "SELECT vector_init('h', 'e', 'type=FLOAT16,dimension=8,distance=L1')",
    ]
    xh = po.convert(rng.standard_normal((20, 8)).astype(np.float32) * 3, po.F16)
    for i in range(20):
        s.append(["INSERT INTO h(id, e) VALUES (?, ?)", [i + 1, blob(xh[i])]])
    s += ["SELECT vector_quantize('h', 'e')", "SELECT rowid1, rowid2, counter, hex(data) FROM vector0_h_e",
          "CREATE TABLE w (k INTEGER PRIMARY KEY, e BLOB) WITHOUT ROWID",
          "SELECT vector_init('w', 'e', 'type=INT8,dimension=4')",
          "CREATE TABLE w2 (k TEXT PRIMARY KEY, e BLOB) WITHOUT ROWID",
          "SELECT vector_init('w2', 'e', 'type=INT8,dimension=4')",
          ["INSERT INTO w VALUES (5, ?)", [blob(np.array([1, -2, 3, 4], dtype=np.int8))]],
          ["INSERT INTO w VALUES (2, ?)", [blob(np.array([9, 9, -9, 0], dtype=np.int8))]],
          "SELECT vector_quantize('w', 'e')", "SELECT rowid1, rowid2, counter, hex(data) FROM vector0_w_e",
          ["INSERT INTO t(id, e) VALUES (999, ?)", [blob(np.zeros(3, dtype=np.float32))]],
          "SELECT vector_quantize('t', 'e')",
          ]
    return s


def scan_script():
    """scans through SQL (needs the GPU for our build; the reference runs them on the CPU)"""
    rng = np.random.Generator(np.random.PCG64(77))
    s = ["CREATE TABLE t (id INTEGER PRIMARY KEY, e BLOB)", "SELECT vector_init('t', 'e', 'type=FLOAT32,dimension=24')"]
    x = rng.standard_normal((600, 24)).astype(np.float32)
    for i in range(600):
        if i % 97 == 5:
            s.append(["INSERT INTO t(id, e) VALUES (?, NULL)", [2 * i + 1]])
        else:
            s.append(["INSERT INTO t(id, e) VALUES (?, ?)", [2 * i + 1, blob(x[i])]])
    q = x[11] + 0.01
    qj = "[" + ",".join(f"{v:.6f}" for v in q) + "]"
    s += [
        ["SELECT id, distance FROM vector_full_scan('t', 'e', ?, 10)", [blob(q)]],
        [f"SELECT rowid, distance FROM vector_full_scan('t', 'e', '{qj}', 5)", []],
        ["SELECT id, distance FROM vector_full_scan('t', 'e', ?, 0)", [blob(q)]],
        ["SELECT count(*) FROM vector_full_scan('t', 'e', ?, 1000)", [blob(q)]],
        ["SELECT id, distance FROM vector_quantize_scan('t', 'e', ?, 10)", [blob(q)]],
        "SELECT vector_quantize('t', 'e', 'max_memory=8KB')",
        ["SELECT id, distance FROM vector_quantize_scan('t', 'e', ?, 10)", [blob(q)]],
        "SELECT vector_quantize_preload('t', 'e')",
        ["SELECT id, distance FROM vector_quantize_scan('t', 'e', ?, 10)", [blob(q)]],
        ["SELECT id, distance FROM vector_quantize_scan('t', 'e', ?, 25) ORDER BY distance DESC", [blob(q)]],
        ["SELECT t.id, v.distance FROM t JOIN vector_quantize_scan('t', 'e', ?, 7) AS v ON t.id = v.rowid", [blob(q)]],
        ["SELECT id, distance FROM vector_full_scan_stream('t', 'e', ?) LIMIT 6", [blob(q)]],
        ["SELECT count(*), min(distance), max(distance) FROM vector_full_scan_stream('t', 'e', ?)", [blob(q)]],
        ["SELECT id, distance FROM vector_quantize_scan_stream('t', 'e', ?) ORDER BY distance, id LIMIT 6", [blob(q)]],
        ["SELECT id, distance FROM vector_full_scan('t', 'e', '[1,2,3]', 5)", []],
        ["SELECT id, distance FROM vector_full_scan('nope', 'e', ?, 5)", [blob(q)]],
        ["INSERT INTO t(id, e) VALUES (5000, ?)", [blob(q)]],
        ["SELECT id, distance FROM vector_full_scan('t', 'e', ?, 3)", [blob(q)]],
        ["DELETE FROM t WHERE id = 5000", []],
        ["SELECT id, distance FROM vector_full_scan('t', 'e', ?, 3)", [blob(q)]],
    ]
    # metrics x types on small tables
    for vt, tname in [(po.F16, "FLOAT16"), (po.BF16, "FLOATB16"), (po.I8, "INT8"), (po.U8, "UINT8")]:
        for metric in ["L2", "SQUARED_L2", "COSINE", "DOT", "L1"]:
            tb = f"m_{tname}_{metric}".lower()
            s += [f"CREATE TABLE {tb} (id INTEGER PRIMARY KEY, e BLOB)", f"SELECT vector_init('{tb}', 'e', 'type={tname},dimension=16,distance={metric}')"]
            xx = po.convert(rng.standard_normal((200, 16)).astype(np.float32), vt)
            for i in range(200):
                s.append([f"INSERT INTO {tb}(id, e) VALUES (?, ?)", [i + 1, blob(xx[i])]])
            s.append([f"SELECT id, distance FROM vector_full_scan('{tb}', 'e', ?, 8)", [blob(xx[3])]])
            s.append(f"SELECT vector_quantize('{tb}', 'e')")
            s.append([f"SELECT id, distance FROM vector_quantize_scan('{tb}', 'e', ?, 8)", [blob(xx[3])]])
    return s
