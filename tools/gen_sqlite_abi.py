"""Derive the slot index of each sqlite3_api_routines member from an SQLite `sqlite3ext.h`
(public-domain SQLite header; pass its path).  The loadable-extension ABI is positional: the
struct is a table of function pointers whose order never changes (new entries are appended), so
sqlite_vector_b200/csrc/sqlite_abi.h only needs the positions of the functions it calls.

    python tools/gen_sqlite_abi.py /path/to/sqlite3ext.h name1 name2 ...
"""
import re
import sys


def members(path):
    src = open(path).read()
    body = src[src.index("struct sqlite3_api_routines {"):]
    body = body[:body.index("\n};")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    depth = 0
    stmt = ""
    for ch in body[body.index("{") + 1:]:
        stmt += ch
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        elif ch == ";" and depth == 0:
            m = re.search(r"\(\s*\*\s*(\w+)\s*\)\s*\(", stmt)      # first "(*name)(" is the member
            if not m:
                m = re.search(r"(\w+)\s*;\s*$", stmt)
            names.append(m.group(1))
            stmt = ""
    return names


if __name__ == "__main__":
    names = members(sys.argv[1])
    want = sys.argv[2:] or names
    for w in want:
        print(f"    VSQ_{w} = {names.index(w)},")
