#!/bin/bash
bash tools/launcher_check.sh
