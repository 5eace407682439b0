import sys, os
sys.path.insert(0, os.getcwd())
import noble_bls12_381_b200 as bls
from noble_bls12_381_b200 import synth
eng = bls.Engine(0, "noble_bls12_381_b200/programs")
gold = open("tests/golden/pairing_kilic_1000.bin", "rb").read()
for n in [int(a) for a in sys.argv[1:]] or [64]:
    g1, g2 = synth.multiples_wire(n)
    out = eng.pairing_batch(g1, g2, n, True)
    k = min(n, 1000)
    print("n", n, "parity", out[:576*k] == gold[:576*k], "kernel ms", eng.last_kernel_ms(), flush=True)
