"""``python -m marl_classification …`` served by the B200 build."""
from marlclassification_b200.__main__ import main

if __name__ == "__main__":
    main()
