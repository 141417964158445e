#!/bin/bash
# tools/sass.sh <object-basename> <mangled-kernel-substring>  -> /tmp/<basename>.sass.txt (one instruction per line)
o=/root/repo/mediastreamer2_b200/lib/obj/$1.o
fn=$(cuobjdump -elf $o 2>/dev/null | grep -o "_Z[A-Za-z0-9_]*" | grep "$2" | sort -u | head -1)
cuobjdump -sass -fun "$fn" $o | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##' > /tmp/$1.sass.txt
echo "$fn: $(wc -l < /tmp/$1.sass.txt) instructions"
awk '{print $2}' /tmp/$1.sass.txt | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${3:-16}
