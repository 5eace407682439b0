"""What can be checked of the Node.js binding without node / tsc in the image: the N-API shim compiles
against a declaration stub of node_api.h, the natives it exports are exactly the ones the TypeScript wrapper calls, the
wrapper exports every function of the reference's surface, the package keeps the reference's export map."""
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JS = os.path.join(ROOT, "js")


def test_addon_compiles_against_the_node_api_stub():
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(JS, "stub"), "-I", os.path.join(ROOT, "include"),
           os.path.join(JS, "addon.cc")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_addon_exports_match_the_wrapper_calls():
    addon = open(os.path.join(JS, "addon.cc")).read()
    ts = open(os.path.join(JS, "index.ts")).read()
    exported = set(re.findall(r'EXPORT\("([A-Za-z0-9]+)"', addon))
    called = set(re.findall(r"\bnative\.([A-Za-z0-9]+)\(", ts))
    assert called, "no native calls found"
    assert called <= exported, called - exported
    # every C entry point the shim calls is declared in the header
    header = open(os.path.join(ROOT, "include", "bls381_b200.h")).read()
    declared = set(re.findall(r"\b(bls381_[a-z0-9_]+)\s*\(", header))
    used = set(re.findall(r"\b(bls381_[a-z0-9_]+)\b", addon)) - {"bls381_b200"}  # the header's file name
    assert used <= declared, used - declared


def test_wrapper_keeps_the_reference_export_surface():
    ts = open(os.path.join(JS, "index.ts")).read()
    # index.ts:22 (re-exports), 94 (utils), 287 / 466 (PointG1 / PointG2), 715-821 (functions)
    for name in ("Fp", "Fr", "Fp2", "Fp12", "CURVE", "PointG1", "PointG2", "utils"):
        assert re.search(r"export \{[^}]*\b%s\b[^}]*\} from '@noble/bls12-381'" % name, ts), name
    for fn in ("pairing", "getPublicKey", "sign", "verify", "aggregatePublicKeys", "aggregateSignatures", "verifyBatch"):
        assert re.search(r"export (async )?function %s\(" % fn, ts), fn
    # the reference's error messages survive
    for msg in ("No pairings at point of Infinity", "Expected non-empty array", "Expected non-empty messages array",
                "Pubkey count should equal msg count", "must be of prime-order subgroup", "Failed to find a square root",
                "Private key must be 0 < key < CURVE.r", "Expected valid private key"):
        assert msg in ts, msg
    # brackets balance (no tsc here: a cheap guard against truncated edits)
    code = re.sub(r"//[^\n]*", "", ts)
    code = re.sub(r"/\*.*?\*/", "", code, flags=re.S)
    code = re.sub(r"'(?:\\.|[^'\\])*'|`(?:\\.|[^`\\])*`", "''", code)
    for o, c in ("()", "[]", "{}"):
        assert code.count(o) == code.count(c), (o, code.count(o), code.count(c))


def test_package_keeps_the_reference_export_map():
    pkg = json.load(open(os.path.join(JS, "package.json")))
    # package.json:46-57 of the reference: "." and "./math", each with types / import / default
    assert set(pkg["exports"]) == {".", "./math"}
    for k in (".", "./math"):
        assert set(pkg["exports"][k]) == {"types", "import", "default"}
    assert pkg["main"] == "lib/index.js" and pkg["module"] == "lib/esm/index.js" and pkg["types"] == "lib/index.d.ts"
    gyp = json.load(open(os.path.join(JS, "binding.gyp")))
    t = gyp["targets"][0]
    assert t["sources"] == ["addon.cc"] and "../include" in t["include_dirs"] and any("bls381_b200" in l for l in t["libraries"])
    assert os.path.exists(os.path.join(JS, "math.ts"))
