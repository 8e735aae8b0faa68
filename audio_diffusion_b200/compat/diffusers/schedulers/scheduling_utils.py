class SchedulerMixin:  # base class of the reference Mel (mel.py:23,44); no behaviour is used
    pass
