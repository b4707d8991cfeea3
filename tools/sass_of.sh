#!/bin/bash
# prints the SASS of one kernel of vtrace_b200/librender.so (substring match on the mangled name), encodings stripped
#   bash tools/sass_of.sh trace_paths_wave_kernelILb1
cuobjdump -sass vtrace_b200/librender.so | awk -v pat="$1" '/Function :/{on=index($0,pat)>0} on' | grep -v "^\s*/\* 0x" | sed 's#/\* 0x[0-9a-f]* \*/##' | cut -c1-110
