"""Build horton_part_b200/data/expbasis_tables.json from the reference's element tables.

The exponential-basis parameters (orders n, exponents alpha, initial coefficients per element) are
scientific input data that the partitioning must use verbatim (SURVEY.md section 2, row 19:
/root/reference/src/horton_part/data/gauss.json, slater.json).  This script re-keys them into one
file  {family: {Z: {"orders": [...], "exponents": [...], "initials": [...]}}}  and is the
provenance record; run it in the build container (the reference tree is not on the GPU box).
"""
import json
import pathlib

SRC = pathlib.Path("/root/reference/src/horton_part/data")
DST = pathlib.Path(__file__).resolve().parents[1] / "horton_part_b200" / "data" / "expbasis_tables.json"

tables = {}
for family in ("gauss", "slater"):
    raw = json.loads((SRC / f"{family}.json").read_text())
    tables[family] = {
        str(int(z)): {"orders": rows[0], "exponents": rows[1], "initials": rows[2]}
        for z, rows in sorted(raw.items(), key=lambda kv: int(kv[0]))
    }
DST.write_text(json.dumps(tables, indent=1) + "\n")
print("wrote", DST)
