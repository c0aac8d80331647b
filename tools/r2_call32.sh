#!/bin/bash
OLD=$PWD/unet_b200/_obj_old/libb2u_oldconv.so
bash tools/ab_step.sh "old_kernel=B2U_LIB=$OLD B2U_NO_STEM_IM2COL=1" "new_default=A=1" "no_im2col=B2U_NO_STEM_IM2COL=1" "no_solo=B2U_CONV_NO_SOLO=1" "stg1=B2U_CONV_MAX_STG=1"
