"""Config schema + loader: the groups of the reference's ``Config`` dataclass that select behaviour of the
env-step path (jobshoplab/types/config_types.py:5-145) and a hydra-free ``load_config``
(jobshoplab/utils/load_config.py:15-46: same arguments; YAML file + dotted ``key=value`` overrides)."""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional, Union

import yaml

from .exceptions import ConfigurationError


@dataclass
class DslRepositoryConfig:
    dir: str = "data/jssp_instances/dsl/default_instance.yaml"


@dataclass
class SpecRepositoryConfig:
    dir: str = "data/jssp_instances/spec_files/ft06"


@dataclass
class CompilerConfig:
    loglevel: str = "warning"
    repo: str = "DslRepository"
    validator: str = "SimpleDSLValidator"
    manipulators: List[str] = field(default_factory=lambda: ["DummyManipulator"])
    dsl_repository: DslRepositoryConfig = field(default_factory=DslRepositoryConfig)
    spec_repository: SpecRepositoryConfig = field(default_factory=SpecRepositoryConfig)


@dataclass
class EnvConfig:
    loglevel: str = "warning"
    observation_factory: str = "BinaryActionObservationFactory"
    reward_factory: str = "BinaryActionJsspReward"
    action_factory: str = "BinaryJobActionFactory"
    render_backend: str = "render_in_dashboard"
    middleware: str = "EventBasedBinaryActionMiddleware"
    max_time_fct: int = 2
    max_action_fct: int = 3


@dataclass
class StateMachineConfig:
    loglevel: str = "warning"
    allow_early_transport: bool = True
    middleware: Optional[str] = None
    action_factory: Optional[str] = None


@dataclass
class EventBasedBinaryActionMiddlewareConfig:
    loglevel: str = "warning"
    truncation_joker: int = 5
    truncation_active: bool = False


@dataclass
class MiddlewareConfig:
    event_based_binary_action_middleware: EventBasedBinaryActionMiddlewareConfig = field(
        default_factory=EventBasedBinaryActionMiddlewareConfig)


@dataclass
class BinaryActionJsspRewardConfig:
    loglevel: str = "warning"
    sparse_bias: float = 1.0
    dense_bias: float = 0.001
    truncation_bias: float = -1.0


@dataclass
class RewardFactoryConfig:
    binary_action_jssp_reward: BinaryActionJsspRewardConfig = field(default_factory=BinaryActionJsspRewardConfig)


@dataclass
class Config:
    title: str = "Dev Environment"
    default_loglevel: str = "warning"
    compiler: CompilerConfig = field(default_factory=CompilerConfig)
    env: EnvConfig = field(default_factory=EnvConfig)
    state_machine: StateMachineConfig = field(default_factory=StateMachineConfig)
    middleware: MiddlewareConfig = field(default_factory=MiddlewareConfig)
    reward_factory: RewardFactoryConfig = field(default_factory=RewardFactoryConfig)
    # groups without influence on the step path (action/observation factory loglevels, render backends)
    extra: dict = field(default_factory=dict)


def _build(cls, data):
    import typing

    kwargs, extra = {}, {}
    hints = typing.get_type_hints(cls)
    for k, v in (data or {}).items():
        if k not in hints:
            extra[k] = v
            continue
        sub = hints[k]
        if dataclasses.is_dataclass(sub) and isinstance(v, dict):
            kwargs[k] = _build(sub, v)
        elif sub is float and isinstance(v, int) and not isinstance(v, bool):
            kwargs[k] = float(v)
        else:
            kwargs[k] = v
    obj = cls(**kwargs)
    if extra and hasattr(obj, "extra"):
        obj.extra = extra
    return obj


def load_config(config_path: Union[str, Path] = "default_config", config_dir: Union[str, Path, None] = None,
                overrides: list[str] | None = None) -> Config:
    """``load_config(config_path="data/config/getting_started_config.yaml")`` like the reference; a bare
    name is looked up as ``<config_dir>/<name>.yaml``."""
    p = Path(config_path)
    if not p.suffix:
        p = Path(config_dir or "data/config") / f"{p.name}.yaml"
    if not p.exists():
        raise ConfigurationError("config_path", str(p))
    data = yaml.safe_load(p.read_text()) or {}
    for ov in overrides or []:
        key, val = ov.split("=", 1)
        cur = data
        parts = key.lstrip("+").split(".")
        for k in parts[:-1]:
            cur = cur.setdefault(k, {})
        cur[parts[-1]] = yaml.safe_load(val)
    return _build(Config, data)
