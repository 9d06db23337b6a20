class LightningCLI:                      # only referenced under `if __name__ == '__main__'` in the trainers
    def __init__(self, *a, **k):
        raise RuntimeError('LightningCLI stand-in: not runnable')
