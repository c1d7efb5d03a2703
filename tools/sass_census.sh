#!/bin/bash
# SASS evidence that the contraction kernels are Blackwell-native (B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,
# UTMALDG/UTMASTG = TMA tile load/store).  Runs without a GPU.
SO=transformertts_b200/libttsb.so
echo "# SASS census of $SO ($(date -u +%F))"
for m in UTCHMMA UTCQMMA LDTM STTM UTMALDG UTMASTG UBLKCP HMMA HGMMA FFMA2 FADD2 REDG.E.ADD.F32x4 BRA.U.ANY; do
  echo "$m $(cuobjdump -sass $SO 2>/dev/null | grep -c "$m")"
done
echo; echo "# per kernel"
cuobjdump -sass $SO 2>/dev/null | awk '/Function : /{name=$3} /UTCHMMA/{mma[name]++} /LDTM/{ld[name]++} /UTMALDG/{tl[name]++} /UTMASTG/{ts[name]++} END{for (k in mma) printf "%s UTCHMMA=%d LDTM=%d UTMALDG=%d UTMASTG=%d\n", k, mma[k], ld[k], tl[k], ts[k]}' | sort
