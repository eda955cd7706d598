"""domain registry (deepxube/factories/domain_factory.py:1-13)."""
from typing import Any, Dict, Tuple

from deepxube_b200.base.factory import Factory
from deepxube_b200.utils.command_line_utils import get_name_args

domain_factory: Factory = Factory("domain")


def get_domain_from_arg(domain: str) -> Tuple[Any, str]:
    """'cube3' | 'npuzzle.4' | 'grid.7d' -> (Domain, domain name)"""
    name, args = get_name_args(domain)
    kwargs: Dict[str, Any] = domain_factory.get_kwargs(name, args)
    return domain_factory.build_class(name, kwargs), name
