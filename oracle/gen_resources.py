#!/usr/bin/env python3
"""Turn reference data files into C headers (xxd -i equivalent) under oracle/_ref/gen/.

Mirrors what the reference's resource compiler does
(lib/mmseqs/cmake/MMseqsResourceCompiler.cmake:26-51): NAME.h defining
`unsigned char NAME_ident[]` and `unsigned int NAME_ident_len`.
Outputs are build artefacts (git-ignored), never committed.
"""
import re, sys, os

def emit(src, outdir):
    name = os.path.basename(src)
    ident = re.sub(r'[^0-9A-Za-z_]', '_', name)
    data = open(src, 'rb').read()
    out = os.path.join(outdir, name + '.h')
    with open(out, 'w') as f:
        f.write('static const unsigned char %s[] = {\n' % ident)
        for i in range(0, len(data), 16):
            f.write('  ' + ', '.join('0x%02x' % b for b in data[i:i+16]) + ',\n')
        f.write('0x00};\n')   # NUL terminator like tostring.sed; len excludes it
        f.write('static const unsigned int %s_len = %d;\n' % (ident, len(data)))

if __name__ == '__main__':
    outdir = sys.argv[1]
    os.makedirs(outdir, exist_ok=True)
    for src in sys.argv[2:]:
        emit(src, outdir)
