#!/bin/bash
# Golden inputs / expected outputs of `merge` and `split`: the reference's own test data (test/test_merge.sh, test/test_split.sh
# compare `slow5tools merge|split` against exactly these files), packed because SLOW5 text is bulky.  Data files only.
# usage: bash tests/golden/make_merge_split_fixtures.sh /root/reference   (writes tests/golden/merge_split_fixtures.tar.xz)
set -e
REF=${1:-/root/reference}; D=$REF/test/data; T=$(mktemp -d); HERE=$(cd "$(dirname "$0")" && pwd)
mkdir -p $T/merge/raw $T/merge/exp $T/split/raw $T/split/exp
cp $D/raw/merge/{rg0,rg0_1,rg0_2_aux_order,rg1_1_new_aux_field,rg0_asic_id_missing,rg0_diff_attr,aux_enum,aux_enum_diff_label,aux_enum_new_label,aux_enum_uint8_t,aux_no_enum}.slow5 \
   $D/raw/merge/{zlib_svb-zd_v0.2.0,zlib_v0.2.0,none_v0.1.0}.blow5 $T/merge/raw/
cp $D/exp/merge/{same_rg,same_rg_aux_order,diff_rg_diff_aux_field,asic_id_missing_expected,same_run_id_different_attribute_values,merged_output_formats,merged_output_enum}.slow5 $T/merge/exp/
cp $D/raw/split/multi_group_blow5s/example_multi_rg_v0.1.0.blow5 $T/split/raw/
cp -r $D/exp/split/expected_group_split_blow5_input $T/split/exp/
chmod -R u+w $T
tar -C $T -cJf $HERE/merge_split_fixtures.tar.xz merge split
rm -rf $T
