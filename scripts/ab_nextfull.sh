# A/B: where the straddle scan tests "the next row is a full row" (cold block vs clamped byte count at the top)
bash scripts/ab_defs.sh 4 "" "RRX_NEXTFULL_CLAMP" "" "RRX_NEXTFULL_CLAMP"
bash scripts/ab_defs.sh 3 "" "RRX_NEXTFULL_CLAMP"
bash scripts/ab_defs.sh 5 "" "RRX_NEXTFULL_CLAMP"
bash scripts/ab_defs.sh 2v "" "RRX_NEXTFULL_CLAMP"
