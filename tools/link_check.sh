#!/bin/sh
# tools/link_check.sh -- does the reference LINK against libscgpu.so with its own hot-path objects removed?
#
# (kept: also src/*.c, src/utils/{entropy,ecc,threading}, the generated ntt_tables.c.)
# Dev-container only (needs /root/reference; nothing is copied into the repository, everything is built under
# /tmp/scgpu_linkcheck).  Compiles every reference source under src/schemes, src/utils/{arith,sampling,crypto}
# EXCEPT the objects libscgpu replaces:
#     utils/arith/ntt.c + the generated ntt_<variant>.c, utils/sampling/{sampling,gaussian_cdf,gaussian_knuth_yao,
#     gaussian_bernoulli}.c, utils/crypto/{prng,prng_get_func,chacha20_csprng,ctr_drbg}.c
# compiles utils/arith/arith.c with its NTT tables and utils_arith_ntt() compiled out (the maintainer's
# `#ifndef HAVE_SCGPU` patch around arith.c:66-396), and links all of it with -lscgpu and -z defs.
# Prints: the symbols the kept objects need from the replaced ones, and every unresolved symbol of the link.
set -e
REF=${REF:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
GEN=$ROOT/oracle/_ref/gen
OUT=/tmp/scgpu_linkcheck
DEFS="-DHAVE_64BIT -DHAVE_128BIT -DUSE_SAFECRYPTO_INTEGER_MP -DUSE_SAFECRYPTO_FLOAT_MP -DHAVE___BUILTIN_CTZ -DSHA3_UNROLLED -DHAVE_AVX -DHAVE_AVX2 -DNTT_NEEDS_7681 -DNTT_NEEDS_12289 -DNTT_NEEDS_8380417 -DNTT_NEEDS_8399873"
INC="-I$GEN -I$REF/include -I$REF/src -I$REF/src/utils/crypto -I$REF/src/utils/arith"
CF="-c -O1 -march=x86-64-v3 -maes -std=gnu11 -fcommon -fPIC -w"
rm -rf $OUT; mkdir -p $OUT/keep $OUT/repl
REPL="$REF/src/utils/arith/ntt.c $GEN/utils/arith/ntt_reference.c $GEN/utils/arith/ntt_barrett.c $GEN/utils/arith/ntt_fp.c $GEN/utils/arith/ntt_avx.c $GEN/utils/arith/ntt_7681.c $GEN/utils/arith/ntt_8380417.c $REF/src/utils/sampling/sampling.c $REF/src/utils/sampling/gaussian_cdf.c $REF/src/utils/sampling/gaussian_knuth_yao.c $REF/src/utils/sampling/gaussian_bernoulli.c $REF/src/utils/crypto/prng.c $REF/src/utils/crypto/prng_get_func.c $REF/src/utils/crypto/chacha20_csprng.c $REF/src/utils/crypto/ctr_drbg.c"
obj() { echo "$2/$(echo "$1" | sed "s#$REF/##; s#$GEN/##; s#/#_#g").o"; }
for f in $REPL; do gcc $CF $DEFS $INC "$f" -o "$(obj "$f" $OUT/repl)"; done
NOTBUILT=""
for f in $(find $REF/src/schemes $REF/src/utils -name '*.c' | grep -v "/unit\|test" | sort) $(ls $REF/src/*.c) $GEN/utils/arith/ntt_tables.c; do
    case " $REPL $REF/src/utils/arith/arith.c " in *" $f "*) continue;; esac
    gcc $CF $DEFS $INC "$f" -o "$(obj "$f" $OUT/keep)" 2>/dev/null || NOTBUILT="$NOTBUILT $(echo $f | sed "s#$REF/##")"
done
awk 'NR==66{print "#ifndef HAVE_SCGPU"} {print} NR==396{print "#endif"}' $REF/src/utils/arith/arith.c > $OUT/arith_patched.c
gcc $CF $DEFS -DHAVE_SCGPU $INC $OUT/arith_patched.c -o $OUT/keep/src_utils_arith_arith.c.o
nm -g --defined-only $OUT/repl/*.o | awk 'NF==3{print $3}' | sort -u > $OUT/repl_defined.txt
nm -g --undefined-only $OUT/keep/*.o | awk '{print $2}' | sort -u > $OUT/keep_undefined.txt
echo "kept objects: $(ls $OUT/keep | wc -l); replaced objects: $(ls $OUT/repl | wc -l)"
echo "not compiled here (need generated config / optional deps):$NOTBUILT"
echo "symbols the kept objects need from the replaced ones:"
comm -12 $OUT/repl_defined.txt $OUT/keep_undefined.txt | tr '\n' ' '; echo
echo "missing from libscgpu.so:"
nm -D --defined-only $ROOT/libsafecrypto_b200/libscgpu.so | awk '{print $3}' | sort -u > $OUT/scgpu_defined.txt
comm -12 $OUT/repl_defined.txt $OUT/keep_undefined.txt | comm -23 - $OUT/scgpu_defined.txt | tr '\n' ' '; echo
echo "link: gcc -shared kept/*.o -lscgpu -Wl,-z,defs"
gcc -shared -fcommon -o $OUT/libsafecrypto_scgpu.so $OUT/keep/*.o -L$ROOT/libsafecrypto_b200 -lscgpu -lm -lpthread -Wl,-z,defs 2>&1 \
    | grep "undefined reference" | sed 's/.*undefined reference to//' | sort | uniq -c || true
echo "link done"
