#!/usr/bin/env python
"""Inject the DDE_MODEL_FN decoration into a model source written for the reference
(plain C functions, e.g. /root/reference/inst/examples/seir.c): every function DEFINITION
at file scope -- a line that starts with a return type and a name followed by `(` -- gets
`DDE_MODEL_FN ` in front.  Nothing else in the file changes.  DDE_MODEL_FN is empty for a C
compiler and `static __device__ __forceinline__` for nvcc (<dde/dde.h>), so the decorated
file still builds against the reference.
usage: tools/decorate_model.py model.c decorated.c"""
import re
import sys

DEFINITION = re.compile(r"^(?!DDE_MODEL_FN)(?:static\s+|inline\s+)*"
                        r"(?:void|double|int|size_t|float|long|unsigned)\b[\w\s\*]*?\b\w+\s*\(")


def decorate(text):
    out = []
    for line in text.splitlines(keepends=True):
        if DEFINITION.match(line) and not line.rstrip().endswith(";"):
            line = "DDE_MODEL_FN " + line
        out.append(line)
    return "".join(out)


if __name__ == "__main__":
    with open(sys.argv[1]) as f:
        text = f.read()
    with open(sys.argv[2], "w") as f:
        f.write(decorate(text))
