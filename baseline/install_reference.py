"""Puts the UNMODIFIED reference package under the git-ignored `baseline/_ref/` so that it travels to the GPU box
(`/root/reference` does not exist there) for `bench.py --impl reference` and the drop-in tests.

    python baseline/install_reference.py

First choice is the contract's offline pip install (`pip install --no-index --no-build-isolation --no-deps
--target baseline/_ref <copy of /root/reference>`).  On this image it fails while generating the package metadata
(setup.py's `setup_requires=['setuptools-markdown']` is not in the wheelhouse), so the fallback does by hand what that
install would have done for this pure-Python distribution (`packages=['hsc']`, setup.py:15): the `hsc/` package
directory is copied byte for byte.  Nothing under baseline/_ref is tracked by git or edited; the Python-2 idioms of the
reference are handled at import time by tests/golden/ref_loader.py.
"""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, '_ref')
SOURCE = os.environ.get('HSC_REFERENCE_SOURCE', '/root/reference')


def installed():
    return os.path.isfile(os.path.join(TARGET, 'hsc', 'modeling.py'))


def up_to_date():
    if not installed():
        return False
    cmp = filecmp.dircmp(os.path.join(SOURCE, 'hsc'), os.path.join(TARGET, 'hsc'), ignore=['__pycache__'])
    return not (cmp.left_only or cmp.diff_files or cmp.funny_files)


def install(verbose=True):
    """Returns 'pip', 'copy' or 'present' (how baseline/_ref/hsc got there); raises if the source tree is missing."""
    if not os.path.isfile(os.path.join(SOURCE, 'hsc', 'modeling.py')):
        if installed():
            return 'present'
        raise RuntimeError('reference source tree not found at %s and baseline/_ref is empty' % SOURCE)
    if up_to_date():
        return 'present'
    how = 'copy'
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, 'reference')
        shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns('.git', '__pycache__'))     # /root/reference is read-only
        cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps', '--find-links',
               '/opt/wheelhouse', '--target', TARGET, src]
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode == 0 and installed():
            how = 'pip'
        else:
            if verbose:
                sys.stderr.write('pip install of the reference failed (%s); copying the hsc/ package instead\n' % (
                    proc.stdout.strip().splitlines()[-1] if proc.stdout.strip() else proc.returncode))
            if os.path.isdir(os.path.join(TARGET, 'hsc')):
                shutil.rmtree(os.path.join(TARGET, 'hsc'))
            os.makedirs(TARGET, exist_ok=True)
            shutil.copytree(os.path.join(SOURCE, 'hsc'), os.path.join(TARGET, 'hsc'), ignore=shutil.ignore_patterns('__pycache__'))
    assert up_to_date()
    return how


if __name__ == '__main__':
    print(install())
