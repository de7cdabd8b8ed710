"""Development tool: prints the constant tables of the host C library's float functions (glibc 2.39, x86-64) that
include/rt_libm.h restates -- read from libm.so.6 at the addresses the functions' code references (objdump -d of the
ifunc-resolved entry points).  Not part of the product."""
import struct as st
import sys

PATH = sys.argv[1] if len(sys.argv) > 1 else "/lib/x86_64-linux-gnu/libm.so.6"
data = open(PATH, "rb").read()
e_phoff = st.unpack_from("<Q", data, 0x20)[0]
e_phentsize, e_phnum = st.unpack_from("<HH", data, 0x36)
segs = []
for i in range(e_phnum):
    p_type, p_flags, p_offset, p_vaddr, p_paddr, p_filesz, p_memsz, p_align = st.unpack_from("<IIQQQQQQ", data, e_phoff + i * e_phentsize)
    if p_type == 1:
        segs.append((p_vaddr, p_offset, p_filesz))


def off(v):
    for va, o, sz in segs:
        if va <= v < va + sz:
            return o + (v - va)
    raise KeyError(hex(v))


def dbl(v, n=1):
    return [st.unpack_from("<d", data, off(v) + 8 * i)[0] for i in range(n)]


def u64(v, n=1):
    return [st.unpack_from("<Q", data, off(v) + 8 * i)[0] for i in range(n)]


def f32(v, n=1):
    return [st.unpack_from("<f", data, off(v) + 4 * i)[0] for i in range(n)]


def u32(v, n=1):
    return [st.unpack_from("<I", data, off(v) + 4 * i)[0] for i in range(n)]


if __name__ == "__main__":
    what = sys.argv[2] if len(sys.argv) > 2 else "expf"
    if what == "expf":
        print("T[32] =", ", ".join("0x%016xull" % x for x in u64(0xb7be0, 32)))
        print("SHIFT", dbl(0xb7d00)[0].hex(), "InvLn2N", dbl(0xb7d08)[0].hex(), "C", [x.hex() for x in dbl(0xb7d10, 3)], "one", dbl(0x98e18)[0].hex())
        print("thresholds", [x.hex() for x in f32(0x8f168, 3)])
    elif what == "raw":
        addr, n, kind = int(sys.argv[3], 16), int(sys.argv[4]), sys.argv[5]
        vals = {"d": dbl, "q": u64, "f": f32, "u": u32}[kind](addr, n)
        print([v.hex() if isinstance(v, float) else hex(v) for v in vals])
