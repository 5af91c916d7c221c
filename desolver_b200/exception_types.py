"""Exception types of the drop-in, same names as /root/reference/desolver/exception_types/exception_types.py:1-13."""


class RecursionError(Exception):
    pass


class FailedIntegration(Exception):
    pass


class FailedToMeetTolerances(Exception):
    pass


__all__ = ["RecursionError", "FailedIntegration", "FailedToMeetTolerances"]
