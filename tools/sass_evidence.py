"""Per-kernel SASS mnemonic counts of the built library (Blackwell evidence): python tools/sass_evidence.py > profiles/sass_evidence_r2.txt
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (TMA), USETMAXREG = setmaxnreg,
SYNCS = mbarrier ops, FFMA2 / FMUL2 = packed f32x2, HMMA = legacy mma.sync (must be 0), DFMA/DMUL/DADD = FP64."""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else 'uecraytracing_b200/libuecrt.so'
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
want = ['UTCHMMA', 'LDTM', 'STTM', 'UTCBAR', 'UBLKCP', 'USETMAXREG', 'SYNCS', 'ELECT', 'FFMA2', 'FMUL2', 'FADD2', 'VIMNMX3', 'VIMNMX', 'SHF', 'HMMA', 'DFMA', 'MUFU', 'REDG', 'ATOMG', 'ATOMS']
cur, counts, sizes = None, collections.OrderedDict(), {}
for l in sass.split('\n'):
    m = re.search(r'Function : (\S+)', l)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r'\(.*\)$', '', cur).replace('void uecrt::', '')
        counts[cur] = collections.Counter(); sizes[cur] = 0
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', l)
    if m and cur:
        sizes[cur] += 1
        op = m.group(1)
        for w in want:
            if op == w or op.startswith(w + '.'):
                counts[cur][w] += 1
                break
print(f"# {lib}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
print(f"{'kernel':58s} {'instr':>6s} " + ' '.join(f'{w:>8s}' for w in want))
tot = collections.Counter()
for k, c in counts.items():
    print(f"{k[:58]:58s} {sizes[k]:6d} " + ' '.join(f'{c[w]:8d}' for w in want))
    tot.update(c)
print(f"{'TOTAL':58s} {sum(sizes.values()):6d} " + ' '.join(f'{tot[w]:8d}' for w in want))
