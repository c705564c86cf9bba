"""Import-only stub (oracle infrastructure): lets /root/reference's src/utils/pylogger.py import."""


def rank_prefixed_message(msg, rank):
    return msg


def rank_zero_only(fn):
    return fn


rank_zero_only.rank = 0
