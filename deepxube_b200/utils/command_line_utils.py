"""`name.args` mini-DSL split (deepxube/utils/command_line_utils.py:10-19)."""
import shlex
import sys
from typing import Optional, Tuple


def print_command() -> None:
    print(" ".join(shlex.quote(a) for a in sys.argv))


def get_name_args(name_args: str) -> Tuple[str, Optional[str]]:
    name, sep, args = name_args.partition(".")
    return name, (args if sep else None)
