"""Loader for tests/golden/reference_outputs.json (outputs of the compiled reference tools, made by
tests/golden/make_golden.py in the build container)."""
import base64
import json
from pathlib import Path

PATH = Path(__file__).resolve().parent / "golden" / "reference_outputs.json"


def load():
    fx = json.loads(PATH.read_text())
    out = {}
    for name, f in fx.items():
        exp = {}
        for k, v in f["expect"].items():
            if k.startswith(("phase_checker", "inbreeding_calculator", "genotype_query", "dosage_calculator")):          # [rc, stdout, stderr text]
                exp[k] = (v[0], base64.b64decode(v[1]), base64.b64decode(v[2]))
                continue
            exp[k] = (v[0], base64.b64decode(v[1])) + tuple(v[2:])
        out[name] = (base64.b64decode(f["input"]), exp)
    return out
