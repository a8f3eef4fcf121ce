"""Builds oracle/_ref/: the REFERENCE's own reduction code (checkm.resultsParser and what it imports), compiled from the
sources where they lie under /root/reference into byte-code files -- the Python counterpart of compiling a C reference into
oracle/_ref/*.so.  No reference source text is copied; oracle/_ref/ is git-ignored and travels to the GPU box like the
other build products.  TEST / BENCH INFRASTRUCTURE: only bench.py's CPU arm (`--impl reference`, cpu_baseline) imports it,
to time "hmmsearch-equivalent search + the reference's ResultsParser" as BASELINE.md prescribes.

    python oracle/build_ref.py          (a no-op where /root/reference does not exist)
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/checkm'
OUT = os.path.join(HERE, '_ref', 'checkm')
MODULES = ['__init__', 'prettytable', 'defaultValues', 'checkmData', 'manifestManager', 'fileEntity', 'common', 'coverage', 'hmmer',
           'hmmerModelParser', 'markerSets', 'resultsParser', 'util/__init__', 'util/pfam', 'util/seqUtils']


def build():
    if not os.path.isdir(REF):
        return False
    for mod in MODULES:
        src = os.path.join(REF, mod + '.py')
        dst = os.path.join(OUT, mod + '.pyc')
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            py_compile.compile(src, cfile=dst, dfile='checkm/' + mod + '.py', doraise=True, quiet=2)
    return True


def available():
    return os.path.exists(os.path.join(OUT, 'resultsParser.pyc'))


def import_reference(data_root):
    """Imports the compiled reference package with CHECKM_DATA_PATH = data_root; returns the modules the CPU arm uses."""
    os.environ['CHECKM_DATA_PATH'] = data_root
    ref_root = os.path.join(HERE, '_ref')
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from checkm.hmmerModelParser import HmmModelParser
        from checkm.markerSets import MarkerSetParser
        from checkm.resultsParser import ResultsParser
    return HmmModelParser, MarkerSetParser, ResultsParser


if __name__ == '__main__':
    print('oracle/_ref built' if build() else '/root/reference not present: nothing built')
