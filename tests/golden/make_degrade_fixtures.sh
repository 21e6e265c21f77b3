#!/bin/bash
# Golden inputs / expected outputs of `degrade`: the reference's own test data (test/test_degrade.sh compares `slow5tools degrade`
# against exactly these files), packed because SLOW5 text is bulky (the text files share their signals, so xz with a large
# dictionary stores them once).  Data files only; the GridION / 5 kHz / RNA004 / P2 Solo pairs are left out for size (same code
# path as the MinION / 4 kHz ones: a different row of the dataset table, covered by the header-detection tests).
# usage: bash tests/golden/make_degrade_fixtures.sh /root/reference   (writes tests/golden/degrade_fixtures.tar.xz)
set -e
REF=${1:-/root/reference}; D=$REF/test/data; T=$(mktemp -d); HERE=$(cd "$(dirname "$0")" && pwd)
mkdir -p $T/degrade/raw $T/degrade/exp
cp $D/raw/degrade/{example2,promr10dna_badhdr,promr10dna_badhdr_sample_freq,promr10dna_badhdr_sample_rate,promr10dna_badhdr_sample_rate2,promr10dna_badrec}.slow5 \
   $D/raw/degrade/{minir10dna,promr10dna4khz,PRPN119035_read1,na12878_prom_merged_r9.4.1_chr22_read1}.blow5 $T/degrade/raw/
cp $D/exp/degrade/example2_b1.slow5 \
   $D/exp/degrade/{example2_b4,minir10dna_b3,promr10dna4khz_b3,PRPN119035_read1_b2,na12878_prom_merged_r9.4.1_chr22_read1_b2}.blow5 $T/degrade/exp/
chmod -R u+w $T
tar -C $T -cf - degrade | xz -9e -T1 > $HERE/degrade_fixtures.tar.xz
rm -rf $T
