"""Name -> class registries and the `<value><suffix>` argument mini-DSL.

Mirrors the behaviour of deepxube/base/factory.py: `Factory.register_class / register_parser / get_kwargs /
build_class / get_type / get_all_class_names` (:124-178), `FactoryAutoBuild` keyed on dataclass field
names + types (:181-233) and `DelimParser` (longest-suffix, case-insensitive token matching, :30-118).
Error behaviour kept: unknown names / parse failures raise ValueError, duplicate registration raises
ValueError, a non-dataclass passed to FactoryAutoBuild.register raises TypeError.
"""
import logging
from abc import ABC, abstractmethod
from dataclasses import fields, is_dataclass
from typing import Any, Callable, Dict, Generic, List, Optional, Tuple, Type, TypeVar, get_type_hints


class Parser(ABC):
    @abstractmethod
    def parse(self, args_str: str) -> Dict[str, Any]:
        ...

    @abstractmethod
    def help(self) -> str:
        ...


class DelimParser(Parser):
    """Tokens split by `delim`; each token is `<value><name>` (or just `<name>` for flags)."""

    def __init__(self) -> None:
        # parse name -> (kwarg name, converter or None for a flag, help, default)
        self._specs: Dict[str, Tuple[str, Optional[Callable[[str], Any]], str, Optional[Any]]] = {}

    @property
    @abstractmethod
    def delim(self) -> str:
        ...

    def add_argument(self, arg_name_parse: str, arg_name: str, value_type: Optional[Callable[[str], Any]], help_msg: str,
                     default: Optional[Any] = None) -> None:
        key = arg_name_parse.lower()
        assert key and arg_name, "length of argument names must be > 0"
        assert self.delim not in key, f"Cannot have delimiter {self.delim} in {key}"
        assert key not in self._specs, f"{key} already exists"
        self._specs[key] = (arg_name, value_type, help_msg, default)

    def _split_token(self, token: str) -> Tuple[str, str]:
        low = token.lower()
        cands = [name for name in self._specs if low.endswith(name)]
        if not cands:
            raise ValueError(f"Unexpected argument {token!r}")
        name = max(cands, key=len)
        return name, token[: len(token) - len(name)]

    def parse(self, args_str: str) -> Dict[str, Any]:
        kwargs: Dict[str, Any] = {}
        if len(args_str) == 0:
            return kwargs
        for token in args_str.split(self.delim):
            assert len(token) > 0, f"Empty argument in {args_str}. Perhaps problem with delimiters."
            name, value = self._split_token(token)
            arg_name, conv, _, _ = self._specs[name]
            if conv is None:
                assert value == "", f"boolean arguments should not have preceding value ({token})"
                kwargs[arg_name] = True
            else:
                assert value != "", f"non-boolean argument {token} does not have preceding value"
                kwargs[arg_name] = conv(value)
        for arg_name, _, _, default in self._specs.values():
            if arg_name not in kwargs and default is not None:
                kwargs[arg_name] = default
        return kwargs

    def help(self) -> str:
        lines = ["Format <arg_val><arg_name> for non-boolean arguments and <arg_name> for boolean arguments. "
                 "<arg_name> is case invariant.", f"Delimited by {self.delim}."]
        for key, (_, conv, msg, default) in self._specs.items():
            usage = key if conv is None else f"<{getattr(conv, '__name__', str(conv))}>{key}"
            line = f"{usage}: {msg}"
            if default is not None:
                line = f"{line} (Default: {default})"
            lines.append(line)
        return "\n".join(lines)


T = TypeVar("T")


class Factory(Generic[T]):
    def __init__(self, class_type_str: str):
        self._classes: Dict[str, Type[T]] = {}
        self._parsers: Dict[str, Type[Parser]] = {}
        self._kind = class_type_str

    def register_class(self, name: str) -> Callable[[Type[T]], Type[T]]:
        def deco(cls: Type[T]) -> Type[T]:
            if name in self._classes:
                raise ValueError(f"{self._kind.capitalize()} {name!r} already registered")
            self._classes[name] = cls
            return cls
        return deco

    def register_parser(self, name: str) -> Callable[[Type[Parser]], Type[Parser]]:
        def deco(cls: Type[Parser]) -> Type[Parser]:
            if name in self._parsers:
                raise ValueError(f"{self._kind.capitalize()} parser {name!r} already registered")
            self._parsers[name] = cls
            return cls
        return deco

    def get_parser(self, name: str) -> Optional[Parser]:
        cls = self._parsers.get(name)
        return cls() if cls is not None else None

    def get_type(self, name: str) -> Type[T]:
        if name not in self._classes:
            raise ValueError(f"Unknown {self._kind} {name!r}. Available: {sorted(self._classes)}")
        return self._classes[name]

    def get_kwargs(self, name: str, args_str: Optional[str]) -> Dict[str, Any]:
        self.get_type(name)
        parser = self.get_parser(name)
        if parser is None or args_str is None:
            assert args_str is None, f"No parser for {self._kind} {name}, however, args given are {args_str}"
            return {}
        try:
            return parser.parse(args_str)
        except Exception as e:  # noqa: BLE001
            logging.exception(f"Error occurred: {e}")
            raise ValueError(f"Error parsing {args_str} for {self._kind} {name!r}.\nParser help:\n{parser.help()}")

    def build_class(self, name: str, kwargs: Dict[str, Any]) -> T:
        return self.get_type(name)(**kwargs)

    def get_all_class_names(self) -> List[str]:
        return list(self._classes)


class FactoryAutoBuild(Generic[T]):
    """Registry of dataclasses, selected by the (field name, field type) set of the data given to build_class."""

    def __init__(self, class_type_str: str):
        self._classes: Dict[Tuple[Tuple[str, Type], ...], Type[T]] = {}
        self._kind = class_type_str

    def register(self, cls: Type[T]) -> Type[T]:
        if not is_dataclass(cls):
            raise TypeError(f"{cls} must be a dataclass")
        hints = get_type_hints(cls)
        key = tuple(sorted(((f.name, hints[f.name]) for f in fields(cls)), key=lambda kv: kv[0]))
        if key in self._classes:
            raise ValueError(f"{self._kind.capitalize()} {key} already registered")
        self._classes[key] = cls
        return cls

    def get_all_class_names(self) -> List[Tuple[Tuple[str, Type], ...]]:
        return list(self._classes)

    def build_class(self, field_data: Dict[str, Any]) -> T:
        for key, cls in self._classes.items():
            if len(key) != len(field_data):
                continue
            if all((n in field_data) and isinstance(field_data[n], t) for n, t in key):
                return cls(**field_data)
        raise ValueError(f"Could not construct {self._kind} class for {field_data}. Tried {list(self._classes)}")
