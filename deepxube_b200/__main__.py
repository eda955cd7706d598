from deepxube_b200._cli import main

if __name__ == "__main__":
    main()
