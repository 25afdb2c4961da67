"""Repository contracts that need no GPU: the reference arm's JSON line, and that the product never touches the oracle."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it."""
    pat = re.compile(r'^\s*(from|import)\s+oracle\b', re.M)
    offenders = []
    for base in ('diffndm_b200', 'scripts'):
        for root, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith('.py') and pat.search(open(os.path.join(root, f)).read()):
                    offenders.append(os.path.join(root, f))
    assert offenders == []


def test_reference_arm_json_line():
    """`bench.py --impl reference` (the oracle port on the host cores): one JSON line with the contract's keys."""
    env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                        '--cpu-batch', '1'], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'ligands/s' and d['higher_is_better'] is True
    assert d['metric'].startswith('ligands/sec') and d['value'] > 0 and d['n_gpus'] == 1 and d['steps'] == 1
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'ligands/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and d['vs_baseline'] is None


def test_scripts_reject_guidance_without_a_host_reward(tmp_path):
    """The reference-flag scripts stop at argument parsing (before any CUDA work) when guidance has no host reward."""
    for script, extra in (('generate_ligands.py', ['--SPSA', '1']), ('inpaint.py', ['--fix_atoms', 'C1', '--svdd', '1'])):
        r = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', script), '--random_init', '0', '--pdbfile', 'x.pdb',
                            '--ref_ligand', 'A:1', '--outfile', str(tmp_path / 'o.sdf')] + extra,
                           capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert r.returncode == 2 and 'need' in r.stderr and '--reward' in r.stderr, r.stderr[-500:]
