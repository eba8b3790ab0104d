#!/bin/bash
# Build variants of the library that differ in one -D flag (kernel experiments); run tools/k2_time.py on each.
#   tools/k2_variants.sh build "ZPX_YCC_WIDE=0 ZPX_YCC_WIDE=1 A=1,B=2 ..."     (here, no GPU needed)
#   tools/k2_variants.sh run   "ZPX_YCC_WIDE=0 ZPX_YCC_WIDE=1 ..."     (on the GPU box)
set -e
cd "$(dirname "$0")/.."
mode=$1
for v in $2; do
  tag=$(echo "$v" | tr '=,' '__')
  v=$(echo "$v" | sed 's/,/ -D/g')   # several flags: A=1,B=2
  so=zpix_b200/variants/libzpixcuda_$tag.so
  if [ "$mode" = build ]; then
    d=/tmp/zpxv_$tag
    rm -rf $d; mkdir -p $d/zpix_b200 zpix_b200/variants
    cp -r include $d/; cp -r zpix_b200/csrc $d/zpix_b200/; rm -f $d/zpix_b200/csrc/*.o
    sed -i "s#^NVFLAGS := #NVFLAGS := -D$v #" $d/zpix_b200/csrc/Makefile
    make -C $d/zpix_b200/csrc -j8 >/dev/null 2>&1 || make -C $d/zpix_b200/csrc
    cp $d/zpix_b200/libzpixcuda.so $so
    echo "built $so"
  else
    echo "== $v"; ZPX_LIB_PATH=$PWD/$so python ${ZPX_PROBE:-tools/k2_time.py}
  fi
done
