"""Version: derived from ``git describe`` in a checkout, static otherwise.

The reference vendors versioneer for the same purpose (/root/reference/mpi4jax/_version.py,
versioneer.py: git tag -> PEP 440).  This is the 30-line version of that: tags look like
``v0.1.0``; ``v0.1.0-5-gabc1234[-dirty]`` becomes ``0.1.0+5.gabc1234[.dirty]``.
"""

import os
import re
import subprocess

_STATIC = "0.1.0"


def _pep440(describe: str):
    m = re.fullmatch(r"v?(\d+(?:\.\d+)*)(?:-(\d+)-g([0-9a-f]+))?(-dirty)?", describe.strip())
    if not m:
        return None
    tag, dist, sha, dirty = m.groups()
    local = []
    if dist and int(dist) > 0:
        local += [dist, "g" + sha]
    if dirty:
        local.append("dirty")
    return tag + ("+" + ".".join(local) if local else "")


def get_version() -> str:
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if os.environ.get("MPI4JAX_B200_STATIC_VERSION") or not os.path.isdir(os.path.join(root, ".git")):
        return _STATIC
    try:
        out = subprocess.run(["git", "describe", "--tags", "--dirty", "--match", "v[0-9]*"], cwd=root,
                             capture_output=True, text=True, timeout=5)
        if out.returncode == 0:
            return _pep440(out.stdout) or _STATIC
    except (OSError, subprocess.SubprocessError):
        pass
    return _STATIC


__version__ = get_version()
