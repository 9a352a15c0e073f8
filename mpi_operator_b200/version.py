"""Build/version info (reference: pkg/version/version.go:21-45)."""
import platform
import sys

__version__ = "0.1.0"
GIT_SHA = "unknown"
BUILT = "unknown"


def info() -> dict:
    return {
        "version": __version__,
        "gitSHA": GIT_SHA,
        "built": BUILT,
        "python": sys.version.split()[0],
        "platform": f"{platform.system().lower()}/{platform.machine()}",
    }


def print_version_and_exit() -> None:
    i = info()
    print(f"Version: {i['version']}\nGit SHA: {i['gitSHA']}\nBuilt: {i['built']}\n"
          f"Python Version: {i['python']}\nOS/Arch: {i['platform']}")
    raise SystemExit(0)
