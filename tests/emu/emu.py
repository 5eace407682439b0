"""ctypes driver for the CPU emulation of the tower-VM interpreter (TEST INFRASTRUCTURE)."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.environ.get("BLS381_EMU_LIB") or os.path.join(HERE, "_build", "libvm_emu.so")
        src = os.path.join(HERE, "vm_emu.cpp")
        deps = [src] + [os.path.join(ROOT, "noble_bls12_381_b200", "csrc", f) for f in ("vm.cuh", "fp_core.cuh", "fp_core_gen.cuh", "fp_inv.cuh", "swu_g2.cuh", "swu_g2_gen.cuh", "g2_kernels.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            os.makedirs(os.path.dirname(so), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", so, src])
        _LIB = ctypes.CDLL(so)
    return _LIB


def run_program(b, buffers, n_items, nslots=72, policy=2):
    """b: scheduled+allocated Builder.  buffers: {buf_id: (bytearray, stride)} (outputs written in place)."""
    prog, nrec = b.encode()
    consts = b.const_table()
    nbuf = max(buffers) + 1
    bases = (ctypes.c_void_p * nbuf)()
    strides = (ctypes.c_uint32 * nbuf)()
    keep = []
    for i, (data, stride) in buffers.items():
        arr = (ctypes.c_uint8 * len(data)).from_buffer(data)
        keep.append(arr)
        bases[i] = ctypes.addressof(arr)
        strides[i] = stride
    rc = lib().vm_emu_run(prog, b.warps, nrec, consts, len(b.consts), b.nslots, b.nfar, n_items,
                          ctypes.cast(bases, ctypes.POINTER(ctypes.POINTER(ctypes.c_uint8))), strides, nbuf, policy)
    assert rc == 0, rc
    return buffers
