class ContextInjectionMiddleware:
    """The reference injects the record's `correlation_id` header into FastStream's context for the
    handler argument (calfkit/client/middleware.py:9-16).  With the batch worker the header travels
    with each Record (calfkit/broker.py) so there is nothing to inject; the name is kept for imports."""
