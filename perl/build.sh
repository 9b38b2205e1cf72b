#!/bin/sh
# Build perl/lib/auto/PedK/PedK.so (xsubpp + gcc; no MakeMaker tree is left behind).
set -e
here=$(cd "$(dirname "$0")" && pwd)
root=$(dirname "$here")
core=$(perl -MConfig -e 'print "$Config{archlibexp}/CORE"')
typemap=$(perl -MConfig -e 'print "$Config{privlibexp}/ExtUtils/typemap"')
ccflags=$(perl -MConfig -e 'print $Config{ccflags}')
mkdir -p "$here/lib/auto/PedK" "$here/build"
xsubpp -typemap "$typemap" "$here/PedK.xs" > "$here/build/PedK.c"
gcc -shared -fPIC -O2 $ccflags -DVERSION=\"0.01\" -DXS_VERSION=\"0.01\" -I"$core" -I"$root/include" \
    "$here/build/PedK.c" -o "$here/lib/auto/PedK/PedK.so" \
    -L"$root/ped_b200/csrc" -lpedk -Wl,-rpath,"$root/ped_b200/csrc"
