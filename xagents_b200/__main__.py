"""`python -m xagents_b200 <command> <agent> ...` -- the reference's `xagents` console entry point (setup.py entry_points)."""
from .cli import execute

if __name__ == '__main__':
    execute()
