"""Exception names of the reference (jobshoplab/utils/exceptions.py) that the env-step path can raise.

The batched engine reports per-env outcomes as status codes (include/jobshoplab_b200.h JSL_ST_*);
the single-env adapter turns a status back into the exception class the reference would have raised
(STATUS_TO_EXCEPTION below), so `except BufferFullError:` written against the reference keeps working.
"""
from __future__ import annotations


class JobShopException(Exception):
    def __init__(self, message=""):
        self.message = message
        super().__init__(message)


def _simple(name, fmt):
    def __init__(self, *args):
        JobShopException.__init__(self, fmt.format(*args, *([""] * 4)))
        self.args_ = args

    return type(name, (JobShopException,), {"__init__": __init__})


class InvalidValue(JobShopException):
    def __init__(self, key, value=None, message=""):
        super().__init__(f"Invalid value for {key}: {value} {message}")
        self.key, self.value = key, value


class NotImplementedError_(JobShopException):
    """The reference's own `NotImplementedError` (it shadows the builtin, exceptions.py:60)."""

    def __init__(self, message="Sorry not done yet"):
        super().__init__(message)


InvalidKey = _simple("InvalidKey", "Invalid key: {0}")
UnsuccessfulStateMachineResult = _simple("UnsuccessfulStateMachineResult", "Unsuccessful state machine result")
ActionOutOfActionSpace = _simple("ActionOutOfActionSpace", "Action {0} is out of action space {1}")
EnvDone = _simple("EnvDone", "Environment is done; call reset()")
BufferFullError = _simple("BufferFullError", "Buffer {0} is full")
JobNotInBufferError = _simple("JobNotInBufferError", "Job {0} not in buffer {1}")
MissingSpecificationError = _simple("MissingSpecificationError", "{0} must be provided")
UnknownLocationNameError = _simple("UnknownLocationNameError", "Unknown location name: {0}")
InvalidDurationError = _simple("InvalidDurationError", "Duration must be an integer, got {0}")
InvalidDistributionError = _simple("InvalidDistributionError", "{0}")
UnknownDistributionTypeError = _simple("UnknownDistributionTypeError", "Unknown distribution type: {0}")
InvalidTimeBehaviorError = _simple("InvalidTimeBehaviorError", "Invalid time behavior: {0}")
InvalidSetupTimesError = _simple("InvalidSetupTimesError", "No setup times for machine {0}")
InvalidToolUsageError = _simple("InvalidToolUsageError", "No tool usage for job {0}")
InvalidTransportConfig = _simple("InvalidTransportConfig", "{0}")
ComponentAssociationError = _simple("ComponentAssociationError", "{1} {0} cannot be associated")
InvalidOutageTypeError = _simple("InvalidOutageTypeError", "Invalid outage type: {0}")
InvalidDispatchRuleError = _simple("InvalidDispatchRuleError", "Invalid dispatch rule {0}; allowed {1}")
TravelTimeError = _simple("TravelTimeError", "No travel time from {0} to {1}")
TransportConfigError = _simple("TransportConfigError", "Invalid transport config {0}: {1}")
ConfigurationError = _simple("ConfigurationError", "Invalid configuration {0}: {1}")
EngineUnavailable = _simple("EngineUnavailable", "{0}")

# JSL_ST_* -> the class the reference raises out of env.step (None: an outcome, not an exception)
STATUS_TO_EXCEPTION = {
    0: None, 1: None, 2: None, 3: None,
    4: BufferFullError,
    5: UnsuccessfulStateMachineResult,
    6: InvalidValue,
    7: EnvDone,
    8: ActionOutOfActionSpace,
    9: NotImplementedError_,
}
