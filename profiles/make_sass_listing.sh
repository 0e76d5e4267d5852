#!/bin/bash
# Evidence that the hot kernels are Blackwell-native: tcgen05 / TMEM / TMA mnemonics in the SASS of the built objects.
#   bash profiles/make_sass_listing.sh > profiles/r02_gram.sass.txt
cd "$(dirname "$0")/../audio_video_textures_b200/csrc" || exit 1
echo "# cuobjdump -sass build/gram.o  (sm_100a; $(nvcc --version | tail -2 | head -1))"
echo "# lines with tcgen05 (UTC*MMA, UTCBAR, UTCATOMSWS ...), TMEM (LDTM), TMA (UTMALDG), mbarrier (SYNCS) mnemonics, per kernel"
cuobjdump -sass build/gram.o | awk '
/Function :/ {fn=$0; sub(/.*Function : /,"",fn); print ""; print "== " fn; next}
/UTC|UTMA|LDTM|STTM|UBLKCP|SYNCS|USETMAXREG|UCGABAR|CGAERRBAR/ {
    line=$0; gsub(/\/\* 0x[0-9a-f]+ \*\//,"",line); gsub(/^[ \t]+/,"",line); sub(/[ \t]+;[ \t]*$/," ;",line); print "  " line }'
echo
echo "# mnemonic histogram (both kernels)"
cuobjdump -sass build/gram.o | grep -oE "\b(UTC[A-Z0-9.]+|UTMA[A-Z0-9.]+|LDTM[A-Z0-9.x]*|SYNCS[A-Z0-9.]+|UCGABAR[A-Z_.]*)" | sort | uniq -c | sort -rn
