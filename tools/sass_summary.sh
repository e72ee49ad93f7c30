#!/usr/bin/env bash
# Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld),
# UBLKCP (cp.async.bulk, the TMA engine's bulk copy), UTCBAR (tcgen05.commit), SYNCS (mbarrier).  Output: profiles/r2_sass_summary.md
set -euo pipefail
HERE="$(cd "$(dirname "$0")/.." && pwd)"
SO="$HERE/alphazero-multi-game_b200/libaz_b200.so"
OUT="$HERE/profiles/r2_sass_summary.md"
TMP="$(mktemp)"; trap 'rm -f "$TMP"' EXIT
cuobjdump -sass "$SO" > "$TMP"
{
  echo "# SASS summary of libaz_b200.so (cuobjdump -sass, sm_100a) — tools/sass_summary.sh"
  echo
  echo "| kernel | UTCHMMA (tcgen05.mma) | of which .2CTA | LDTM (tcgen05.ld) | UBLKCP (TMA bulk copy) | UTCBAR (tcgen05.commit) | SYNCS (mbarrier) | instructions |"
  echo "|---|---|---|---|---|---|---|---|"
  awk '
    /Function : / { if (name != "") print name "|" mma "|" mma2 "|" ldtm "|" blk "|" bar "|" syn "|" ins; name=$3; mma=0; mma2=0; ldtm=0; blk=0; bar=0; syn=0; ins=0; next }
    /^ +\/\*[0-9a-f]+\*\/ / { ins++ }
    /UTCHMMA/ { mma++ } /UTCHMMA\.2CTA/ { mma2++ } /LDTM/ { ldtm++ } /UBLKCP/ { blk++ } /UTCBAR/ { bar++ } /SYNCS/ { syn++ }
    END { if (name != "") print name "|" mma "|" mma2 "|" ldtm "|" blk "|" bar "|" syn "|" ins }
  ' "$TMP" | while IFS='|' read -r n a b c d e f g; do
      dn="$(echo "$n" | c++filt | sed 's/(anonymous namespace):://g; s/(.*//; s/^void //; s/az::nn:://; s/az:://')"
      echo "| \`$dn\` | $a | $b | $c | $d | $e | $f | $g |"
    done | sort -t'|' -k3 -n -r
  echo
  echo "Totals: UTCHMMA $(grep -c UTCHMMA "$TMP"), of which .2CTA $(grep -c 'UTCHMMA.2CTA' "$TMP"); LDTM $(grep -c LDTM "$TMP"); UBLKCP $(grep -c UBLKCP "$TMP"); UTCBAR $(grep -c UTCBAR "$TMP"); UTMALDG (tensor-map TMA) $(grep -c UTMALDG "$TMP" || true) — operands are fed by 1-D bulk copies of the padded position stream, not by tensor maps."
} > "$OUT"
echo "wrote $OUT"
