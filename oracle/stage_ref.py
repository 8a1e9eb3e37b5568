"""TEST / BENCH INFRASTRUCTURE -- stages the UNMODIFIED reference where it can travel to the GPU box.

The reference (cwfparsonson/ddls) is pure Python: there is nothing to compile.  "Building" it for the CPU arm of bench.py
(``--impl reference``) means copying its package, byte for byte, from /root/reference/ddls to oracle/_ref/ddls.  oracle/_ref/
is git-ignored (reference sources never enter the history) but not gpurun-ignored, so it ships with the snapshot like the
built .so files.  Run by ``__graft_entry__.build()`` whenever /root/reference is present:

    python oracle/stage_ref.py
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get('DDLS_REFERENCE_SRC', '/root/reference')
DST = os.path.join(HERE, '_ref')


def stage(force=False):
    """Returns the staged root (oracle/_ref) or None when the reference is not available here."""
    src_pkg = os.path.join(SRC, 'ddls')
    if not os.path.isdir(src_pkg):
        return DST if os.path.isdir(os.path.join(DST, 'ddls')) else None
    dst_pkg = os.path.join(DST, 'ddls')
    if os.path.isdir(dst_pkg) and not force:
        cmp = filecmp.dircmp(src_pkg, dst_pkg, ignore=['__pycache__'])
        if not (cmp.left_only or cmp.right_only or cmp.diff_files):
            return DST
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(src_pkg, dst_pkg, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    with open(os.path.join(DST, 'STAGED_FROM.txt'), 'w') as f:
        f.write(f'{SRC}/ddls copied verbatim by oracle/stage_ref.py (unmodified reference, cwfparsonson/ddls @ 9e0b5ba)\n')
    return DST


if __name__ == '__main__':
    out = stage(force='--force' in sys.argv)
    print(out or 'reference not available')
