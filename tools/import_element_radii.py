"""Build horton_part_b200/data/element_radii.json from the reference's constants table.

The Becke scheme must use the reference's per-element radii verbatim
(/root/reference/src/horton_part/data/constants.yaml: ``radius_becke`` = Slater covalent radii where
available, else Cordero; ``radius_covalent`` = Cordero), in angstrom, indexed by atomic number.
Scientific input data, re-keyed; this script is the provenance record (build container only).
"""
import json
import pathlib

import yaml

SRC = pathlib.Path("/root/reference/src/horton_part/data/constants.yaml")
DST = pathlib.Path(__file__).resolve().parents[1] / "horton_part_b200" / "data" / "element_radii.json"

doc = yaml.safe_load(SRC.read_text())
out = {
    "unit": "angstrom",
    "radius_becke": {str(z): v for z, v in enumerate(doc["radius_becke"]) if v is not None},
    "radius_covalent": {str(z): v for z, v in enumerate(doc["radius_covalent"]) if v is not None},
}
DST.write_text(json.dumps(out, indent=1) + "\n")
print("wrote", DST)
