#!/bin/bash
# SASS evidence for the tensor-core / TMA claims of DESIGN.md s3 (run here, no GPU needed):
#   bash profiles/sass_excerpts.sh > profiles/r02_sass_excerpts.txt
LIB=xpsi_b200/libxpsi_b200.so
echo "cuobjdump -sass $LIB  ($(date -u +%F), nvcc $(nvcc --version | grep -o 'release [0-9.]*'))"
echo
echo "instruction counts per kernel (DMMA = fp64 tensor core mma.sync m8n8k4, UTMALDG = TMA tensor load, STL = local-memory store):"
cuobjdump -sass $LIB 2>/dev/null | awk '
  /Function :/ { f=$3 }
  /DMMA/ { d[f]++ } /UTMALDG/ { t[f]++ } /STL/ { s[f]++ } /SYNCS/ { m[f]++ }
  /^[ \t]+\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f][0-9a-f]*\*\// { n[f]++ }
  END { for (k in n) if (d[k] || t[k]) printf "%6d instr  DMMA %3d  UTMALDG %d  mbarrier(SYNCS) %d  STL %3d  %s\n", n[k], d[k], t[k], m[k], s[k], k }' | sort -k9 | c++filt | sed 's/(xb::AzinvArgs.*//; s/(xb::FoldArgs.*//'
echo
for fn in '_ZN2xb16k_azinv_flux_mmaILi2ELi0ELi100EEEvNS_9AzinvArgsE14CUtensorMap_stS2_' '_ZN2xb10k_fold_mmaENS_8FoldArgsE'; do
  echo "== $(echo $fn | c++filt | sed 's/(.*//'): first tensor-core / TMA instructions"
  cuobjdump -sass -fun "$fn" $LIB 2>/dev/null | grep -E "UTMALDG|DMMA|SYNCS" | head -8 | sed 's/^\s*//; s/\s*\/\* 0x[0-9a-f]* \*\///'
  echo
done
