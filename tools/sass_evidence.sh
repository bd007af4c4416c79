#!/bin/bash
# Counts the Blackwell-specific SASS mnemonics per kernel of libdformer_b200.so (no GPU needed): UTC*MMA = tcgen05.mma,
# LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP = TMA, FFMA2 = packed fp32 FMA, HMMA = warp-level mma.sync (only the 49-row attention core, gaa_mma.cu: see its header for why not tcgen05).
#   tools/sass_evidence.sh > profiles/r02_sass_mnemonics.txt
LIB=${1:-dformer_b200/libdformer_b200.so}
echo "cuobjdump -sass $LIB  (sm_100a; built by 'make'; $(nvcc --version | tail -1))"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { fn=$3; names[fn]=1; next }
  { for (i=1;i<=NF;i++) { t=$i
      if (t ~ /^UTC[A-Z]*MMA/) a[fn,"UTCMMA"]++
      else if (t ~ /^LDTM/) a[fn,"LDTM"]++
      else if (t ~ /^STTM/) a[fn,"STTM"]++
      else if (t ~ /^UTMALDG/) a[fn,"UTMALDG"]++
      else if (t ~ /^UTMASTG/) a[fn,"UTMASTG"]++
      else if (t ~ /^UBLKCP/) a[fn,"UBLKCP"]++
      else if (t ~ /^FFMA2/) a[fn,"FFMA2"]++
      else if (t ~ /^HMMA/) a[fn,"HMMA"]++
      else if (t ~ /^SYNCS/) a[fn,"SYNCS"]++
      else if (t ~ /^LDGSTS/) a[fn,"LDGSTS"]++
      else if (t ~ /^RED\./ || t ~ /^REDG/) a[fn,"RED"]++
      else if (t ~ /^MUFU/) a[fn,"MUFU"]++ } }
  END { printf "%-9s %-6s %-6s %-8s %-8s %-7s %-7s %-6s %-7s %-6s %-6s %-6s kernel\n","UTC*MMA","LDTM","STTM","UTMALDG","UTMASTG","UBLKCP","FFMA2","HMMA","SYNCS","LDGSTS","RED","MUFU"
        for (fn in names) { s=0; split("UTCMMA LDTM STTM UTMALDG UTMASTG UBLKCP FFMA2 HMMA",k," "); for (j in k) s+=a[fn,k[j]]
          if (s>0) printf "%-9d %-6d %-6d %-8d %-8d %-7d %-7d %-6d %-7d %-6d %-6d %-6d %s\n", a[fn,"UTCMMA"],a[fn,"LDTM"],a[fn,"STTM"],a[fn,"UTMALDG"],a[fn,"UTMASTG"],a[fn,"UBLKCP"],a[fn,"FFMA2"],a[fn,"HMMA"],a[fn,"SYNCS"],a[fn,"LDGSTS"],a[fn,"RED"],a[fn,"MUFU"], fn } }' | c++filt | cut -c1-230 | (read -r hdr; echo "$hdr"; read -r hdr2; echo "$hdr2"; sort -k13)
