#!/usr/bin/env bash
# Per-object histogram of the SASS opcodes that prove which hardware paths the kernels use (cuobjdump -sass on the
# objects the library is linked from): DMMA = FP64 tensor cores (mma.sync m8n8k4 f64; tcgen05 has no f64 kind),
# UTMALDG / UTMASTG = TMA tensor loads / stores, UBLKCP = TMA bulk copies, SYNCS = mbarrier operations,
# LDGSTS = cp.async, UTCMMA / tcgen05 would be the 5th-generation tensor path (absent: no f64 kind).
set -u
cd "$(dirname "$0")/.."
printf "%-26s %8s %8s %8s %8s %8s %8s %8s %8s\n" object DMMA DFMA UTMALDG UTMASTG UBLKCP SYNCS LDGSTS UTCMMA
for o in build/obj/*.o; do
  s=$(cuobjdump -sass "$o" 2>/dev/null)
  c() { printf "%s" "$s" | grep -c -E "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ +)?$1[ .]" ; }
  printf "%-26s %8s %8s %8s %8s %8s %8s %8s %8s\n" "$(basename "$o")" "$(c DMMA)" "$(c DFMA)" "$(c UTMALDG)" "$(c UTMASTG)" "$(c UBLKCP)" "$(c SYNCS)" "$(c LDGSTS)" "$(c UTCMMA)"
done
