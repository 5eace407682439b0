#!/usr/bin/env python3
"""Extract the reference's own golden vectors for the pairing / verify / sign hot path into small,
self-contained fixtures under tests/golden/ (the GPU box has no /root/reference).

Run in the build container:   python tests/golden/make_golden.py

Sources (all under /root/reference/test, read-only; DATA only, no reference source code is copied):
  go_pairing_vectors/pairing.json      -> pairing_kilic_1000.bin   (1000 x 576 B, already re-ordered to
                                          noble's Fp12.toBytes() order; deterministic.test.ts:34-46)
  pairing.test.ts:46-96                -> pairing_kats.json        (e(G1,G2) and finalExponentiate KATs)
  bls12-381-g2-test-vectors.txt        -> sign_g2_vectors.txt      (559 priv:msg:sig lines)
  zkcrypto/*.dat                       -> zkcrypto_*.dat           (4 x 1000 encodings of i*G)
  hashToCurve.test.ts                  -> hash_to_curve.json       (xmd + G2 RO/NU/kilic vectors)
"""
import json
import os
import re
import shutil

REF = "/root/reference/test"
OUT = os.path.dirname(os.path.abspath(__file__))


def kilic_to_noble(hexstr: str) -> bytes:
    parts = re.findall(".{96}", hexstr)
    assert len(parts) == 12
    return bytes.fromhex("".join(reversed(parts)))  # deterministic.test.ts:41


def js_strings(expr: str) -> str:
    """Concatenate every '...' literal of a JS string expression ('a' + 'b')."""
    return "".join(re.findall(r"'([^']*)'", expr))


def split_objects(block: str):
    depth = 0
    start = None
    for i, ch in enumerate(block):
        if ch == "{":
            if depth == 0:
                start = i
            depth += 1
        elif ch == "}":
            depth -= 1
            if depth == 0:
                yield block[start + 1 : i]


def array_block(src: str, name: str) -> str:
    m = re.search(r"const %s = \[" % re.escape(name), src)
    assert m, name
    i = m.end()
    depth = 1
    j = i
    while depth:
        c = src[j]
        if c == "[":
            depth += 1
        elif c == "]":
            depth -= 1
        j += 1
    return src[i : j - 1]


def parse_vectors(src: str, name: str):
    out = []
    for obj in split_objects(array_block(src, name)):
        m_msg = re.search(r"msg:(.*?)(?=\n\s*(?:len|expected):)", obj, re.S)
        m_len = re.search(r"len:\s*(0x[0-9a-fA-F]+|\d+)", obj)
        m_exp = re.search(r"expected:(.*)$", obj, re.S)
        v = {"msg": js_strings(m_msg.group(1)), "expected": js_strings(m_exp.group(1))}
        if m_len:
            v["len"] = int(m_len.group(1), 0)
        out.append(v)
    return out


def main():
    # 1. kilic pairings
    vecs = json.load(open(f"{REF}/go_pairing_vectors/pairing.json"))
    assert len(vecs) == 1000
    with open(f"{OUT}/pairing_kilic_1000.bin", "wb") as f:
        for v in vecs:
            f.write(kilic_to_noble(v))

    # 2. pairing.test.ts KATs
    src = open(f"{REF}/pairing.test.ts").read()
    m = re.search(r"vectors from https://github.com/zkcrypto/pairing.*?fromBigTwelve\(\[(.*?)\]\)", src, re.S)
    e_g1_g2 = [int(x, 16) for x in re.findall(r"0x([0-9a-f]+)n", m.group(1))]
    m = re.search(r"finalExponentiate is correct.*?fromBigTwelve\(\[(.*?)\]\).*?fromBigTwelve\(\[(.*?)\]\)", src, re.S)
    fe_in = [int(x) for x in re.findall(r"(\d{50,})n", m.group(1))]
    fe_out = [int(x, 16) for x in re.findall(r"0x([0-9a-f]+)n", m.group(2))]
    assert len(e_g1_g2) == len(fe_in) == len(fe_out) == 12
    json.dump(
        {
            "source": "test/pairing.test.ts:46-96",
            "e_g1_g2": [hex(x) for x in e_g1_g2],
            "final_exp_in": [hex(x) for x in fe_in],
            "final_exp_out": [hex(x) for x in fe_out],
        },
        open(f"{OUT}/pairing_kats.json", "w"),
        indent=1,
    )

    # 3. sign vectors, scalar vectors, zkcrypto
    shutil.copyfile(f"{REF}/bls12-381-g2-test-vectors.txt", f"{OUT}/sign_g2_vectors.txt")
    for n in ("g1_compressed", "g1_uncompressed", "g2_compressed", "g2_uncompressed"):
        shutil.copyfile(f"{REF}/zkcrypto/{n}_valid_test_vectors.dat", f"{OUT}/zkcrypto_{n}.dat")
    for p in os.listdir(OUT):
        os.chmod(f"{OUT}/{p}", 0o644)

    # 4. hash-to-curve vectors (xmd, G2, and -- SURVEY section 8f-4 -- the G1 suites)
    src = open(f"{REF}/hashToCurve.test.ts").read()
    long_dst = js_strings(re.search(r"const LONG_DST =(.*?);", src, re.S).group(1))
    h2c = {
        "source": "test/hashToCurve.test.ts",
        "xmd_sha256": {"dst": "QUUX-V01-CS02-with-expander-SHA256-128", "vectors": parse_vectors(src, "VECTORS")},
        "xmd_sha256_long_dst": {"dst": long_dst, "vectors": parse_vectors(src, "VECTORS_BIG")},
        "xmd_sha512": {"dst": "QUUX-V01-CS02-with-expander-SHA512-256", "vectors": parse_vectors(src, "VECTORS_SHA512")},
        "g2_kilic_ro": {"dst": "BLS12381G2_XMD:SHA-256_SSWU_RO_TESTGEN", "vectors": parse_vectors(src, "VECTORS_G2")},
        "g2_rfc_ro": {"dst": "QUUX-V01-CS02-with-BLS12381G2_XMD:SHA-256_SSWU_RO_", "vectors": parse_vectors(src, "VECTORS_G2_RO")},
        "g2_rfc_nu": {"dst": "QUUX-V01-CS02-with-BLS12381G2_XMD:SHA-256_SSWU_NU_", "vectors": parse_vectors(src, "VECTORS_G2_NU")},
        "g2_kilic_nu": {"dst": "BLS12381G2_XMD:SHA-256_SSWU_NU_TESTGEN", "vectors": parse_vectors(src, "VECTORS_ENCODE_G2")},
        "g1_kilic_ro": {"dst": "BLS12381G1_XMD:SHA-256_SSWU_RO_TESTGEN", "vectors": parse_vectors(src, "VECTORS_G1")},
        "g1_rfc_ro": {"dst": "QUUX-V01-CS02-with-BLS12381G1_XMD:SHA-256_SSWU_RO_", "vectors": parse_vectors(src, "VECTORS_G1_RO")},
        "g1_rfc_nu": {"dst": "QUUX-V01-CS02-with-BLS12381G1_XMD:SHA-256_SSWU_NU_", "vectors": parse_vectors(src, "VECTORS_G1_NU")},
        "g1_kilic_nu": {"dst": "BLS12381G1_XMD:SHA-256_SSWU_NU_TESTGEN", "vectors": parse_vectors(src, "VECTORS_ENCODE_G1")},
    }
    json.dump(h2c, open(f"{OUT}/hash_to_curve.json", "w"), indent=1)
    print({k: len(v["vectors"]) for k, v in h2c.items() if isinstance(v, dict)})


if __name__ == "__main__":
    main()
