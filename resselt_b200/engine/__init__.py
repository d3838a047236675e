from .module import EngineModule, ParamSpec, random_state_dict
from .plan import INPUT, OUTPUT, Plan, PlanBuilder, Ref

__all__ = ['EngineModule', 'ParamSpec', 'random_state_dict', 'INPUT', 'OUTPUT', 'Plan', 'PlanBuilder', 'Ref']
