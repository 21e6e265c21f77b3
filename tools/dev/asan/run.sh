#!/bin/bash
# AddressSanitizer + UBSan + LeakSanitizer over the HOST side of the library and the CLI (no device needed): the four host sources
# and the CLI are built with -fsanitize=address,undefined against stubs of the GPU entry points (gpu_stubs.cpp: every codec call
# fails with S5B_ERR_DEVICE, so only uncompressed / text paths run), then the C programs of tests/test_boundary.py and a few CLI
# commands of the CPU tests are run.  usage: bash tools/dev/asan/run.sh [/root/reference]      (prints one line per run)
set -e
REF=${1:-/root/reference}; R=$(cd "$(dirname "$0")/../../.." && pwd); H=$R/slow5tools_b200/csrc/host; T=$(mktemp -d); cd $T
sed "s#/root/repo/include#$R/include#" $R/tools/dev/asan/gpu_stubs.cpp > stubs.cpp
FL="-g -O1 -std=c++11 -fsanitize=address,undefined -fno-omit-frame-pointer"
g++ $FL -fPIC -shared $H/blow5_io.cpp $H/s5b_file_api.cpp $H/press_api.cpp $H/index_main.cpp stubs.cpp -o libslow5b200.so
g++ $FL $H/view_main.cpp $H/get_main.cpp $H/merge_split_main.cpp $H/degrade_main.cpp -o cli -L . -lslow5b200 -lpthread -Wl,-rpath,$T
python3 - "$R" <<'PY'
import re, sys
src = open(sys.argv[1] + "/tests/test_boundary.py").read()
for k in ("AUX_PROG", "GET_PROG", "INTRO_PROG", "WRITE_PROG"):
    open(k + ".c", "w").write(re.search(k + r' = r"""(.*?)"""', src, re.S).group(1))
PY
for p in AUX_PROG GET_PROG INTRO_PROG WRITE_PROG; do gcc -g -O1 -w -fsanitize=address,undefined -I $R/include/compat $p.c -o $p -L . -lslow5b200 -Wl,-rpath,$T; done
export ASAN_OPTIONS=detect_leaks=1 S5B_ORDERLY_EXIT=1
D=$REF/test/data; EX=$REF/slow5lib/examples/example.slow5
run() { name=$1; shift; "$@" > $name.out 2> $name.err && rc=0 || rc=$?; echo "$name rc=$rc sanitizer_reports=$(grep -c 'ERROR: AddressSanitizer\|ERROR: LeakSanitizer\|runtime error' $name.err || true)"; }
run view ./cli view $EX -o ex.blow5 -c none -s none
run view_enum ./cli view $D/raw/merge/aux_enum.slow5 -o enum.blow5 -c none -s none
run aux ./AUX_PROG $R/tests/golden/fixtures/exp_1_lossless.blow5
run get ./GET_PROG ex.blow5 $(grep -v '^[#@]' $EX | cut -f1 | head -2) no-such-read
run intro ./INTRO_PROG enum.blow5 end_reason
run write ./WRITE_PROG w.blow5
run write_text ./WRITE_PROG w.slow5
gcc -g -O1 -w -fsanitize=address,undefined -I $R/include/compat $REF/slow5lib/examples/append.c -o append -L . -lslow5b200 -Wl,-rpath,$T
cp w.blow5 test.blow5
run append ./append
run index ./cli index ex.blow5
run cli_get ./cli get ex.blow5 $(grep -v '^[#@]' $EX | cut -f1 | head -1) --to slow5
run demux ./cli split -x $D/raw/split/demux9/barcode_summary.txt $D/raw/split/demux10/example2_0_multi.slow5 -d o1 --to slow5 --demux-rid rid --demux-code code
run demux_blow5 ./cli split -x $D/raw/split/demux11/onemissing.txt $D/raw/split/demux10/example2_0_multi.slow5 -d o2 --to blow5 -c none -s none --demux-rid rid --demux-code code -m rest
run merge ./cli merge $D/raw/merge/rg0.slow5 $D/raw/merge/rg1_1_new_aux_field.slow5 -o m.slow5
run split_groups ./cli split -g $D/raw/split/multi_group_slow5s/rg.slow5 -d o3 --to slow5
run degrade_badhdr ./cli degrade $D/raw/degrade/promr10dna_badhdr.slow5
cd /; rm -rf $T
