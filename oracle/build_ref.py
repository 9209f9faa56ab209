#!/usr/bin/env python
"""Recipe that ships the UNMODIFIED reference beside the oracle (test / benchmark infrastructure).

    python oracle/build_ref.py            # /root/reference/skge/*.py -> oracle/_ref/skge/

The reference is pure Python, so "building" it is copying its package directory, byte for byte,
into ``oracle/_ref/`` (git-ignored: the sources never enter this repository's history; NOT
gpurun-ignored: the copy travels to the GPU box with the snapshot, where ``bench.py --impl
reference`` and the ``cpu_baseline`` leg time the real ``HolEEval.positions`` /
``TransEEval.positions`` of skge/base.py:913-1031 on the host cores).  ``MANIFEST.json`` records the
sha256 of every file so a reader can check that nothing was edited.  Nothing in the product path
(scikit-kge_b200/) reads ``oracle/``; only tests, ``__graft_entry__`` and bench.py's CPU legs do.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')


def build(src=None, quiet=False):
    src = src or os.environ.get('SKGE_REFERENCE', '/root/reference')
    pkg = os.path.join(src, 'skge')
    if not os.path.isdir(pkg):
        if not quiet:
            print('reference not found at %s: keeping the existing oracle/_ref (if any)' % src)
        return None
    out = os.path.join(DST, 'skge')
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(out)
    manifest = {}
    for name in sorted(os.listdir(pkg)):
        if not name.endswith('.py'):
            continue
        data = open(os.path.join(pkg, name), 'rb').read()
        with open(os.path.join(out, name), 'wb') as f:
            f.write(data)
        manifest['skge/' + name] = hashlib.sha256(data).hexdigest()
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': src, 'files': manifest}, f, indent=1, sort_keys=True)
    if not quiet:
        print('oracle/_ref: %d files copied unmodified from %s' % (len(manifest), pkg))
    return DST


if __name__ == '__main__':
    sys.exit(0 if build() else 1)
