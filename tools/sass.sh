#!/bin/bash
# usage: tools/sass.sh <kernel-name-substring> > out.sass   (plain SASS of one kernel from libisscabac.so)
cuobjdump -sass /root/repo/isscabac_b200/libisscabac.so 2>/dev/null | awk '/Function : /{name=$3} {print name "\t" $0}' | grep "$1" | cut -f2 | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed 's#^\s*/\*\([0-9a-f]*\)\*/\s*#\1 #; s#/\*.*##'
