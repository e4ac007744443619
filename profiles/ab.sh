for l in "$@"; do DVF_LIB_NAME=$l python profiles/kernel_time.py $KT_ARGS 2>&1 | tail -1; done
