#!/bin/bash
# Opcode evidence of the shipped library: per kernel, how many tcgen05 MMAs (UTCHMMA, .2CTA = cta_group::2), TMEM loads
# (LDTM), TMA loads (UTMALDG), multicast commits (UTCBAR), mbarrier ops (SYNCS), reductions (RED/REDG) the SASS holds.
# usage: tools/sass_summary.sh > profiles/r2_sass_opcodes.md
SO=video-to-action-release_b200/libv2a_b200.so
echo "# SASS opcode counts per kernel of \`$SO\` (cuobjdump -sass, sm_100a)"
echo
echo "| kernel | instructions | UTCHMMA | of which .2CTA | LDTM | UTMALDG | UTCBAR | SYNCS | REDG/RED | STG | LDG |"
echo "|---|---|---|---|---|---|---|---|---|---|---|"
cuobjdump -sass $SO 2>/dev/null | awk '
/Function : /{ if (name != "") print_row(); name=$3; n=0; mma=0; mma2=0; ldtm=0; tma=0; bar=0; syncs=0; red=0; stg=0; ldg=0 }
/^[ \t]+\/\*[0-9a-f]+\*\//{ n++; if ($0 ~ /UTCHMMA/) {mma++; if ($0 ~ /2CTA/) mma2++}
  if ($0 ~ /LDTM/) ldtm++; if ($0 ~ /UTMALDG/) tma++; if ($0 ~ /UTCBAR/) bar++; if ($0 ~ /SYNCS/) syncs++;
  if ($0 ~ /RED/) red++; if ($0 ~ / STG/) stg++; if ($0 ~ / LDG/) ldg++ }
function print_row() { if (mma + ldtm + tma > 0 || name ~ /prep|stencil|smallm|fingerprint|cfg_step|ddim/) printf("| `%s` | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d |\n", name, n, mma, mma2, ldtm, tma, bar, syncs, red, stg, ldg) }
END{ print_row() }' | sed 's/_ZN3v2a//' | c++filt 2>/dev/null
