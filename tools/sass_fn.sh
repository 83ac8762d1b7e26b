#!/bin/bash
# usage: tools/sass_fn.sh <object> <mangled-name> : plain SASS listing (address + instruction) of one kernel
cuobjdump -sass -fun "$2" "$1" 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's#^\s+/\*([0-9a-f]+)\*/\s+#\1 #; s#\s*/\* 0x[0-9a-f]+ \*/##; s#\s+;#;#'
