#!/bin/bash
# Golden inputs / expected outputs of `split -x` (demultiplexing): the reference's own test data (test/test_split.sh compares
# `slow5tools split -x ...` against exactly these directories).  Data files only.
# usage: bash tests/golden/make_demux_fixtures.sh /root/reference   (writes tests/golden/demux_fixtures.tar.xz)
set -e
REF=${1:-/root/reference}; D=$REF/test/data; T=$(mktemp -d); HERE=$(cd "$(dirname "$0")" && pwd)
mkdir -p $T/demux/raw $T/demux/exp
cp -r $D/raw/split/demux* $T/demux/raw/
cp -r $D/exp/split/demux* $T/demux/exp/
chmod -R u+w $T
tar -C $T -cf - demux | xz -9e -T1 > $HERE/demux_fixtures.tar.xz
rm -rf $T
