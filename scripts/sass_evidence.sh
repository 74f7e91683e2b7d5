# usage: bash scripts/sass_evidence.sh [out]  -- SASS mnemonic counts of the shipped library (regenerate after every kernel change)
OUT=${1:-profiles/r2_sass_evidence.txt}
LIB=openmpl_b200/libmpl_b200.so
{
  echo "# SASS evidence of $LIB (sha256 $(sha256sum $LIB | cut -c1-16), git $(git rev-parse --short HEAD), $(date -u +%F))"
  echo "# cuobjdump -sass | grep -c <mnemonic>;  UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG / UTMAREDG / UTMAPF = TMA, HMMA = mma.sync (SPT, head)"
  cuobjdump -sass $LIB > /tmp/_sass.txt
  for m in UTCHMMA UTCHMMA.2CTA LDTM UTMALDG UTMASTG UTMAREDG UTMAPF UTCBAR SYNCS HMMA MUFU.EX2 MUFU.TANH; do
    printf "%-14s %s\n" $m $(grep -c "$m" /tmp/_sass.txt)
  done
  echo "# kernels containing UTCHMMA:"
  awk '/Function :/{f=$3} /UTCHMMA/{c[f]++} END{for(k in c) print c[k], k}' /tmp/_sass.txt | sort -k2 | sed 's/_ZN3mpl[0-9]*_GLOBAL__N__[0-9a-f_]*gemm_tcgen05_cu_[0-9a-f]*//' | head -40
} > $OUT
cat $OUT | head -20
