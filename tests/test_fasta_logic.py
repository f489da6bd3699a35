"""CPU: the bit-parallel FASTA record rules of the device parser (dashing_b200/csrc/fasta_logic.h — plain host/device functions
over byte-class masks and lane ballots) against a byte-by-byte state machine, on random text and every incoming state."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fasta_logic_matches_bytewise_rules(tmp_path):
    exe = str(tmp_path / "fasta_logic_test")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "fasta_logic_test.cpp")])
    r = subprocess.run([exe, "20000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
