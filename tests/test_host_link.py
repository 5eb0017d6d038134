"""The C++ host mirror must define everything its header declares (a shared library links with undefined members as
long as nothing references them): compile a program that references the whole surface and link it against the library."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_header_surface_links(tmp_path, hr):
    pkg = os.path.join(ROOT, "hanamaru_renderer_b200")
    exe = str(tmp_path / "host_linkcheck")
    cmd = ["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(pkg, "csrc", "host"),
           os.path.join(ROOT, "tests", "cpp", "host_linkcheck.cpp"), "-o", exe, "-L", pkg, "-lhanamaru_host", "-Wl,-rpath," + pkg]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert run.returncode == 0, run.stderr[-2000:]
