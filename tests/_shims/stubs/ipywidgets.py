class Button:
    def __init__(self, *a, **k):
        pass
