"""TEST INFRASTRUCTURE ONLY (oracle shim): reference utils/util.py:156 only needs
pytz.timezone(zone) as a tzinfo for a log-directory timestamp."""
import datetime


def timezone(zone):
    return datetime.timezone.utc
