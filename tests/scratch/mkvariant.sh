#!/bin/bash
# usage: mkvariant.sh name [extra nvcc -D flags...]   -> tests/scratch/variants/lib_<name>.so
name=$1; shift
root=$(cd "$(dirname "$0")/../.." && pwd)
mkdir -p $root/tests/scratch/variants
make -C $root/jacobi_pd_b200/csrc OUT=$root/tests/scratch/variants/lib_$name.so OBJ=$root/tests/scratch/variants/obj_$name EXTRA="$*" 2>&1 | grep -E "error|Error"
echo built $name
