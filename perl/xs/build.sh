#!/bin/bash
# Builds the Perl XS shim perl/xs/auto/Chr3D/XS/XS.so against include/chr3d.h and libchr3d.so.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
CORE="$(perl -MConfig -e 'print "$Config{archlibexp}/CORE"')"
TYPEMAP="$(perl -MConfig -e 'print "$Config{privlibexp}/ExtUtils/typemap"')"
CC="$(test -x /usr/bin/gcc && echo /usr/bin/gcc || echo gcc)"
mkdir -p "$HERE/auto/Chr3D/XS"
xsubpp -typemap "$TYPEMAP" "$HERE/Chr3D.xs" > "$HERE/Chr3D.c"
$CC -shared -fPIC -O2 $(perl -MConfig -e 'print $Config{ccflags}') -I"$CORE" -I"$ROOT/include" \
    -o "$HERE/auto/Chr3D/XS/XS.so" "$HERE/Chr3D.c" \
    -L"$ROOT/chromosome3d_b200" -l:libchr3d.so -Wl,-rpath,"\$ORIGIN/../../../../../chromosome3d_b200"
echo "built $HERE/auto/Chr3D/XS/XS.so"
