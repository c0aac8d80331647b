"""Blackwell-specific opcodes per kernel of libb2u.so (cuobjdump -sass), written to stdout.
usage: python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "unet_b200", "libb2u.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
OPS = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "UTCBAR", "UCGABAR", "SYNCS", "R2UR", "HMMA"]
kern, counts, order = None, {}, []
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        order.append(kern)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and kern:
        op = m.group(1)
        counts[kern]["instructions"] += 1
        if op.startswith("UTCHMMA"):
            counts[kern]["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
        else:
            for o in OPS[2:]:
                if op.startswith(o):
                    counts[kern][o] += 1
print("cuobjdump -sass unet_b200/libb2u.so (sm_100a): Blackwell-specific opcodes per kernel.  UTCHMMA = tcgen05.mma (.2CTA = cta_group::2),\n"
      "UTMALDG / UTMASTG = TMA tensor load / store, UTMAPF = tensor-map prefetch, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,\n"
      "UCGABAR = cluster barrier, SYNCS = mbarrier operations, R2UR = register -> uniform-register moves.  Kernels without any\n"
      "of them (the streaming kernels) are summarised in the last line.  No HMMA (mma.sync) anywhere in the library.\n")
tot, plain = collections.Counter(), 0
for k in order:
    c = counts[k]
    tot.update({o: c[o] for o in OPS})
    if not any(c[o] for o in OPS[:9]):
        plain += 1
        continue
    name = re.sub(r"\(.*", "", demangle(k))
    print(name)
    print(f"      {c['instructions']} instructions:  " + "  ".join(f"{o} {c[o]}" for o in OPS[:10] if c[o]))
print(f"\n{plain} streaming kernels (BatchNorm, pooling, PixelShuffle/concat, im2col, loss, optimizers, casts, stitching, attention glue): none of the opcodes above")
print("library totals: " + "  ".join(f"{o} {tot[o]}" for o in OPS))
