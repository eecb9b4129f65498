"""Import shim for the reference's optional UPN proposal detector (``detect_tools/upn``), which is OUT OF SCOPE of
this engine: ``inference.py:3`` imports the name without using it."""


class UPNWrapper:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("the UPN proposal detector is not part of the fo1-b200 engine: pass proposal boxes in")
