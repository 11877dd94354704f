"""The pure-host entry points of the library (octile packer, threaded batch
packer, node reorderings) under AddressSanitizer + UBSan and under
ThreadSanitizer: tests/native/pack_fuzz.cpp drives them with a few hundred
ragged random graphs (single nodes, no edges, isolated nodes, self loops,
parallel edges, sizes around the 8- and 32-boundaries, feature pools), checks
that bad input is refused, that the reorderings are permutations and that the
batch packer writes the per-graph packer's bytes."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = [os.path.join(ROOT, 'tests', 'native', 'pack_fuzz.cpp'),
           os.path.join(ROOT, 'graphdot_b200', 'csrc', 'gdb_pack.cpp')]
INCLUDES = ['-I', os.path.join(ROOT, 'include'),
            '-I', os.path.join(ROOT, 'graphdot_b200', 'csrc')]


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
@pytest.mark.parametrize('name, flags, rounds', [
    ('asan_ubsan', ['-fsanitize=address,undefined',
                    '-fno-sanitize-recover=undefined'], 300),
    ('tsan', ['-fsanitize=thread'], 40)])
def test_host_entry_points_under_sanitizers(tmp_path, name, flags, rounds):
    exe = str(tmp_path / f'pack_fuzz_{name}')
    build = subprocess.run(['g++', '-std=c++17', '-O1', '-g', *flags,
                            *INCLUDES, *SOURCES, '-o', exe, '-lpthread'],
                           capture_output=True, text=True)
    if build.returncode != 0 and 'sanitizer' in build.stderr.lower() \
            and 'cannot find' in build.stderr.lower():
        pytest.skip('sanitizer runtime not installed')
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe, str(rounds)], capture_output=True, text=True,
                         timeout=600)
    assert run.returncode == 0, (run.stdout + run.stderr)[-3000:]
    assert 'pack_fuzz ok' in run.stdout
    assert 'Sanitizer' not in run.stderr
