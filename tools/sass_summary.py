"""Per-kernel count of the Blackwell-only SASS mnemonics in the in-tree library:
  UTCIMMA  = tcgen05.mma kind::i8      LDTM    = tcgen05.ld (TMEM -> registers)
  UTMALDG  = TMA tensor load           UTMASTG = TMA tensor store      UTMAPF = TMA prefetch
  UTCBAR   = tcgen05.commit            SYNCS   = mbarrier ops
Usage: python tools/sass_summary.py [lib.so] > profiles/r02_sass_tc_i8.txt   (no GPU needed)"""
import collections
import hashlib
import re
import subprocess
import sys
from pathlib import Path

lib = Path(sys.argv[1] if len(sys.argv) > 1 else "mixdq_b200/libmixdq_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCIMMA|UTCHMMA|UTCQMMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UTCBAR|UTCCP|SYNCS|"
                 r"IMMA|HMMA|IDP4A|IDP|ELECT|ACQBULK|UBLKCP)\b")
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    if cur is not None:
        m = pat.search(line)
        if m:
            counts[cur][m.group(1)] += 1
        if re.search(r"/\*[0-9a-f]{4,}\*/\s+[A-Z@]", line):
            counts[cur]["_instr"] += 1
print(f"# {lib}  sha256 {hashlib.sha256(lib.read_bytes()).hexdigest()[:16]}  (cuobjdump -sass, sm_100a)")
print("# instructions | mnemonic counts")
for k, c in counts.items():
    body = " ".join(f"{n}={v}" for n, v in sorted(c.items()) if n != "_instr")
    print(f"{k:90s} {c['_instr']:6d} | {body}")
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("# total:", " ".join(f"{n}={v}" for n, v in sorted(tot.items()) if n != "_instr"))
